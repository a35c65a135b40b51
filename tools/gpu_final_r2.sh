#!/bin/bash
# round-2 final evidence pass on one GPU
mkdir -p gpurun_out
cp MEASURED_PEAKS.json gpurun_out/ 2>/dev/null
bash tools/gpu_profile.sh r2b
timeout -k 10 600 ncu --set full --clock-control none -k regex:"route_kernel|slab_splat" -s 4 -c 2 -f -o gpurun_out/r2b_prof_routed python tools/debug/prof_routed.py > gpurun_out/r2b_routed.log 2>&1
timeout -k 10 900 python tools/config_sweep.py 3 4 5 > gpurun_out/r2b_config_sweep.json 2> gpurun_out/r2b_config_sweep.err; echo "sweep rc=$?"; tail -c 300 gpurun_out/r2b_config_sweep.err
timeout -k 10 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err; echo "ref rc=$?"
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 --n-points 1024 --cpu-sample-pairs 100000 --votes network > gpurun_out/r2b_config1_cpu.json 2>> gpurun_out/r2b_bench_ref.err; echo "config1 rc=$?"
ls -la gpurun_out | tail -12
