#!/bin/bash
# evidence pass: launch list of the one-call pipeline + full ncu of the top kernels (never a bench value)
TAG=${1:-r2}
mkdir -p gpurun_out
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/${TAG}_ncu_b.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:'encode_sample_tc|vote_private|backvote_bins|point_encode_tc_kernel|knn_kernel|survivor_stats' -s 18 -c 6 -f -o gpurun_out/${TAG}_prof_kernels python bench.py --streams 1 --steps 2 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
