#!/bin/bash
# A/B of run-time switches on one box: bench (no variants, no CPU arm) per setting -> gpurun_out/ab_<name>.json
mkdir -p gpurun_out
run() {   # name, env...
    local name=$1; shift
    env "$@" python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-variants > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
    python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/ab_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["ms_per_step"], 4), {k: round(v["avg_ms"], 4) for k, v in d["kernels"].items()})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
run base CPPF_TC_WAIT_HINT=0
run hint1k CPPF_TC_WAIT_HINT=1000
run hint100k CPPF_TC_WAIT_HINT=100000
run hint10m CPPF_TC_WAIT_HINT=10000000
run base2 CPPF_TC_WAIT_HINT=0
