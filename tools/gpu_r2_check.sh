#!/bin/bash
# round-2 GPU pass: all GPU tests (no -x: every failure is wanted), smoke, one bench line
mkdir -p gpurun_out
cp MEASURED_PEAKS.json gpurun_out/ 2>/dev/null
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout -k 10 1500 python -m pytest tests -m gpu -q -s --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?" | tee -a gpurun_out/bench1.err
tail -5 gpurun_out/bench1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench1.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"])
    print("kernels", {k: round(v["avg_ms"], 4) for k, v in d["kernels"].items()})
    print("roofline", {k: v for k, v in d["roofline"].items() if k != "note"})
    print("sampled", {k: v for k, v in (d.get("variant_sampled_100k") or {}).items() if k not in ("note",)})
    print("trained", {k: v for k, v in (d.get("variant_trained_network") or {}).items() if k not in ("note",)})
    g = d.get("vs_gpu_reference") or {}
    for k, v in g.items():
        if isinstance(v, dict):
            print("gpuref", k, {kk: vv for kk, vv in v.items() if kk != "stages"})
            for s, r in v.get("stages", {}).items():
                print("   ", s, r)
    print("cpu", {k: v for k, v in (d.get("cpu_baseline") or {}).items() if k != "sample"})
except Exception as e:
    print("bench parse failed", e)
PY
