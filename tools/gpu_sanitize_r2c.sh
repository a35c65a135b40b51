#!/bin/bash
# round 2, final build: compute-sanitizer over the tensor-core SPRIN kernel (memcheck + racecheck on every point-encoder test
# except the N = 4096 cases) and memcheck over the batched entry with the single-memset workspace
mkdir -p gpurun_out
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_point_encoder.py tests/test_gpu_batch.py -m gpu -x -q -k "not 4096" > gpurun_out/r2c_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_memcheck.log
tail -5 gpurun_out/r2c_memcheck.log
timeout -k 10 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 30 python -m pytest tests/test_gpu_point_encoder.py -m gpu -x -q -k "not 4096 and not knn" > gpurun_out/r2c_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_racecheck.log
tail -5 gpurun_out/r2c_racecheck.log
