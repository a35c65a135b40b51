#!/bin/bash
# A/B of library variants (tools/build_variants.py) on one box: bench (no variants, no CPU arm) per library
# -> gpurun_out/ab_<name>.json.  Usage: tools/gpu_ab.sh base name1 name2 ...   ("base" = the regular library)
mkdir -p gpurun_out
for name in "$@"; do
    lib=""
    [ "$name" != base ] && [ "$name" != base2 ] && lib="$PWD/cppf_b200/_variants/libcppf_$name.so"
    CPPF_B200_LIB=$lib python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-variants > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
    python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/ab_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    k = d["kernels"]
    print(f"{sys.argv[1]:10s} step {d['ms_per_step']:.4f}  enc {k['encode_sample']['avg_ms']:.4f}  vote {k['vote']['avg_ms']:.4f}  "
          f"bv {k['backvote']['avg_ms']:.4f}  pe {k['point_encoder']['avg_ms']:.4f}  stats {k['stats']['avg_ms']:.4f}  rot {k['rot_hist']['avg_ms']:.4f}")
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
