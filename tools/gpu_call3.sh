#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_fused.py -m gpu -x -q > gpurun_out/pytest_fused.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_fused.log
timeout -k 10 400 python bench.py --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "rc=$?" >> gpurun_out/bench_tc.err
