#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 120 tools/_bin/dsmem_atomics_bench > gpurun_out/dsmem_atomics.json 2> gpurun_out/dsmem_atomics.err; echo "rc=$?" >> gpurun_out/dsmem_atomics.err
timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout -k 10 900 python tools/config_sweep.py > gpurun_out/config_sweep.json 2> gpurun_out/config_sweep.err; echo "rc=$?" >> gpurun_out/config_sweep.err
ls -la gpurun_out
