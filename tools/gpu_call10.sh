#!/bin/bash
# one-call pipeline (cppf_pose_fused) + vote v3: parity tests, smoke, bench, full ncu of the vote kernel
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout -k 10 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:'vote_private' -s 3 -c 1 -o gpurun_out/prof_vote3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
