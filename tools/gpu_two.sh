#!/bin/bash
# final build on 2 GPUs of one box (run under gpurun --gpus 2): every GPU test (including the NCCL world-2 row-split test
# that a 1-GPU box skips), the bench line at 2 GPUs, the row split of one N = 16384 object
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -k 10 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout -k 10 400 $TR --nproc-per-node 2 --master-port 29582 bench.py --gpus 2 --steps 20 --warmup 3 --no-variants > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err; echo "bench2 rc=$?"
timeout -k 10 300 $TR --nproc-per-node 2 --master-port 29702 tools/bench_rowsplit.py --n-points 16384 --steps 3 --warmup 1 > gpurun_out/r2c_rowsplit_n2.json 2> gpurun_out/r2c_rowsplit_n2.err; echo "rowsplit2 rc=$?"; tail -c 500 gpurun_out/r2c_rowsplit_n2.json
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c_bench_n2.json").read().strip().splitlines()[-1])
    print("bench n=2 value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "n_gpus", d["n_gpus"])
except Exception as e:
    print("bench2 parse failed", e)
PY
