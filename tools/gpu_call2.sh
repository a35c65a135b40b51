#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/_bin/tc_probe > gpurun_out/tc_probe.log 2>&1; echo "rc=$?" >> gpurun_out/tc_probe.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
