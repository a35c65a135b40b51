#!/bin/bash
# One gpurun call: GPU parity suite, smoke, the bench line of both arms.  Usage: gpurun --timeout 1800 -- 'bash tools/gpu_check.sh'
# (tools/gpu_quick.sh = tests + bench without the CPU arm; tools/gpu_profile.sh = launch list + ncu captures;
#  tools/gpu_memcheck.sh = compute-sanitizer; tools/gpu_scale4.sh = 4-GPU torchrun bench)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout -k 10 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
