#!/bin/bash
# BASELINE config 5: stress sweep N in {2k, 4k, 8k, 16k} x GPUs in {1, 2, 4, 8}, weak scaling (one dense object per rank per
# step), each cell one bench.py line with its self-measured roofline block (run under gpurun --gpus 8)
mkdir -p gpurun_out
: > gpurun_out/r2b_config5_grid.jsonl
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2048 4096 8192 16384; do
  steps=10; warm=3
  [ $n -ge 16384 ] && steps=4 && warm=2
  for g in 1 2 4 8; do
    if [ $g -eq 1 ]; then
      timeout -k 10 600 python bench.py --gpus 1 --n-points $n --steps $steps --warmup $warm --no-variants --no-cpu-baseline >> gpurun_out/r2b_config5_grid.jsonl 2> gpurun_out/r2b_config5_n${n}_g${g}.err
    else
      timeout -k 10 600 $TR --nproc-per-node $g --master-port $((29800+g)) bench.py --gpus $g --n-points $n --steps $steps --warmup $warm --no-variants --no-cpu-baseline >> gpurun_out/r2b_config5_grid.jsonl 2> gpurun_out/r2b_config5_n${n}_g${g}.err
    fi
    echo "n=$n g=$g rc=$?"
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/r2b_config5_grid.jsonl"):
    l = l.strip()
    if not l.startswith("{"): continue
    d = json.loads(l)
    r = d.get("roofline") or {}
    print(d["config"]["n_points"], d["n_gpus"], round(d["ms_per_step"], 3), round(d["value"] / 1e9, 3), "Gpairs/s", "e2e", round(d["e2e"]["value"] / 1e9, 3), "roof", r.get("kernel"), round(r.get("frac") or 0, 3))
PY
