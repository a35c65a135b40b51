#!/bin/bash
# 8-GPU torchrun bench (both arms), as the driver launches it
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_L8.txt
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "rc=$?" >> gpurun_out/bench_n8.err
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29582 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench_ref_n8.json 2>> gpurun_out/bench_n8.err
