#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout -k 10 400 python bench.py --no-cpu-baseline > gpurun_out/bench_tc2.json 2> gpurun_out/bench_tc2.err; echo "rc=$?" >> gpurun_out/bench_tc2.err
