#!/bin/bash
# final build, 8 GPUs of one box (gpurun --gpus 8): the bench line and BASELINE config 4 as one strong-scaling point
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -k 10 400 $TR --nproc-per-node 8 --master-port 29588 bench.py --gpus 8 --steps 20 --warmup 3 --no-variants > gpurun_out/r2c_bench_n8.json 2> gpurun_out/r2c_bench_n8.err; echo "bench8 rc=$?"
timeout -k 10 300 $TR --nproc-per-node 8 --master-port 29608 tools/scale_config4.py > gpurun_out/r2c_config4_n8.json 2> gpurun_out/r2c_config4_n8.err; echo "config4 n=8 rc=$?"; tail -c 500 gpurun_out/r2c_config4_n8.json
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c_bench_n8.json").read().strip().splitlines()[-1])
    print("bench n=8 value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"])
except Exception as e:
    print("bench8 parse failed", e)
PY
