#!/bin/bash
# Round-end evidence in one gpurun call: parity suite, smoke, both bench arms, launch list and full ncu capture of the top kernels
bash tools/gpu_check.sh
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-variants > gpurun_out/ncu_b.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:'encode_sample_tc|vote_private|backvote_bins|point_encode_kernel|knn_kernel|survivor_stats' -s 12 -c 6 -o gpurun_out/prof_kernels python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-variants > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; ls -la gpurun_out | tail -8
