#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:'encode_sample_tc|vote_private' -s 2 -c 2 -o gpurun_out/prof_r1b python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r1b.log 2>&1
