#!/bin/bash
# evidence pass: full bench (+ reference arm), launch list of the one-call pipeline, full ncu of the top kernels
mkdir -p gpurun_out
cp MEASURED_PEAKS.json gpurun_out/ 2>/dev/null
timeout -k 10 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?" >> gpurun_out/bench_full.err
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_full.err
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-variants > gpurun_out/ncu_b.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:'encode_sample_tc|vote_private|backvote_bins|point_encode_kernel|knn_kernel|survivor_stats' -s 12 -c 6 -o gpurun_out/prof_kernels python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-variants > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
