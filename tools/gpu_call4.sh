#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 120 tools/_bin/tc_probe > gpurun_out/tc_probe.log 2>&1; echo "rc=$?" >> gpurun_out/tc_probe.log
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:'encode_sample_tc' -s 1 -c 1 -o gpurun_out/prof_tc python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_tc.log 2>&1
