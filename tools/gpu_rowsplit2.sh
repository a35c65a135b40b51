#!/bin/bash
# 2-GPU box: row-split tests (lockstep + NCCL) and the strong-scaling bench of one dense object
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_rowsplit.py -m gpu -x -q > gpurun_out/pytest_rowsplit.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_rowsplit.log
tail -5 gpurun_out/pytest_rowsplit.log
for n in 8192 16384; do
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      tools/bench_rowsplit.py --n-points $n --steps 4 --warmup 2 > gpurun_out/rowsplit_n${n}_g2.json 2> gpurun_out/rowsplit_n${n}_g2.err
  tail -1 gpurun_out/rowsplit_n${n}_g2.json
done
