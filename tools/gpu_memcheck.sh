#!/bin/bash
# compute-sanitizer memcheck over the GPU parity suite subset that exercises the new kernels
mkdir -p gpurun_out
timeout -k 10 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_fused.py tests/test_gpu_configs.py tests/test_gpu_scene.py tests/test_gpu_preprocess.py -m gpu -x -q -k "not overflow_flush" > gpurun_out/memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/memcheck.log
tail -30 gpurun_out/memcheck.log
