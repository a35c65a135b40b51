#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:'route_kernel|slab_splat' -s 22 -c 4 -o gpurun_out/prof_routed python tools/config_sweep.py 3 > gpurun_out/ncu_routed.log 2>&1
ls -la gpurun_out
