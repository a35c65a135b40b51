"""Compile-time variants of libcppf_b200.so for A/B runs on the GPU box (tools/gpu_ab.sh):
    python tools/build_variants.py name=file.cu:-DKNOB=1[,-DOTHER=2] ...
recompiles `file.cu` with the extra flags, links it with the other objects of the regular build
(cppf_b200/_obj/*.o) into cppf_b200/_variants/libcppf_<name>.so; select one with CPPF_B200_LIB=<path>."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cppf_b200 import build as B          # noqa: E402


def main():
    B.build()
    out_dir = os.path.join(B.HERE, "_variants")
    os.makedirs(out_dir, exist_ok=True)
    for spec in sys.argv[1:]:
        name, rest = spec.split("=", 1)
        fname, flags = rest.split(":", 1)
        flags = [f for f in flags.split(",") if f]
        src = os.path.join(B.CSRC, fname)
        obj = os.path.join(B.OBJ, f"variant_{name}_{fname[:-3]}.o")
        r = subprocess.run([B.NVCC, *B.FLAGS, *flags, "-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit(f"nvcc failed for variant {name}:\n{r.stdout}{r.stderr}")
        regs = [l for l in (r.stdout + r.stderr).splitlines() if "registers" in l or "spill" in l]
        others = [os.path.join(B.OBJ, f) for f in sorted(os.listdir(B.OBJ))
                  if f.endswith(".o") and not f.startswith("variant_") and f != fname[:-3] + ".o"]
        lib = os.path.join(out_dir, f"libcppf_{name}.so")
        r = subprocess.run([B.NVCC, "-shared", "-o", lib, obj, *others, "-lcudart", "-Xlinker", "--no-undefined"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit(f"link failed for variant {name}:\n{r.stdout}{r.stderr}")
        print(name, "->", lib)
        for l in regs[-6:]:
            print("   ", l.strip()[:160])


if __name__ == "__main__":
    main()
