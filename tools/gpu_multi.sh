#!/bin/bash
# multi-GPU evidence on one box (run under gpurun --gpus 8): config 4 strong scaling, row split of one N=16384 object,
# the NCCL world-2 row-split parity test, and bench.py at 8 GPUs
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_smi_L.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do
  timeout -k 10 300 $TR --nproc-per-node $n --master-port $((29600+n)) tools/scale_config4.py > gpurun_out/r2_config4_n$n.json 2> gpurun_out/r2_config4_n$n.err; echo "config4 n=$n rc=$?"; tail -c 600 gpurun_out/r2_config4_n$n.json
done
for n in 2 4 8; do
  timeout -k 10 300 $TR --nproc-per-node $n --master-port $((29700+n)) tools/bench_rowsplit.py --n-points 16384 --steps 3 --warmup 1 > gpurun_out/r2_rowsplit_n$n.json 2> gpurun_out/r2_rowsplit_n$n.err; echo "rowsplit n=$n rc=$?"; tail -c 700 gpurun_out/r2_rowsplit_n$n.json
done
timeout -k 10 300 python -m pytest tests/test_gpu_rowsplit.py -q 2>&1 | tail -3
timeout -k 10 600 $TR --nproc-per-node 8 --master-port 29581 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench8 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_n8.json").read().strip().splitlines()[-1])
    print("bench n=8 value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"])
except Exception as e:
    print("bench8 parse failed", e)
PY
