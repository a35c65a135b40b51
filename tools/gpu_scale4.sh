#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_L4.txt
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "rc=$?" >> gpurun_out/bench_n4.err
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/bench_ref_n4.json 2>> gpurun_out/bench_n4.err
