"""TEST INFRASTRUCTURE (oracle) -- builds the *reference's own* voting kernels.

The reference ships its four CUDA kernels as Python string literals handed to
``cupy.RawKernel`` (``/root/reference/models/voting.py:4,70,115,150``).  CuPy is not
in this image, so ``import models.voting`` fails; instead this recipe

1. parses ``models/voting.py`` with ``ast`` and pulls out each RawKernel's source
   string and entry-point name (the file is never imported or copied),
2. makes a *temporary* copy of ``models/include/helper_math.cuh`` whose hard-coded
   ``/usr/local/cuda-10.2`` include (line 27) points at ``<cuda_runtime.h>``; the
   ``findpeak`` string includes the non-existent ``models/src/helper_math.cuh``
   (voting.py:151) so the temp tree mirrors the header there too,
3. compiles every string for the GPU  -> ``oracle/_ref/ref_<name>.cubin``
   (``nvcc -arch=sm_100a``, no fast-math, fmad on -- CuPy/NVRTC defaults), and
4. compiles the same strings for the CPU -> ``oracle/_ref/libref_voting_cpu.so``
   (g++ -fopenmp through ``oracle/ref_cpu_shim.h`` plus tiny launch loops written
   here).

Outputs go only into ``oracle/_ref/`` (git-ignored, travels to the GPU box).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline/reference
arms may load them.  Needs ``/root/reference``; on the GPU box the prebuilt files
are used as-is.
"""
from __future__ import annotations

import ast
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("CPPF_REFERENCE", "/root/reference")

# launch loops for the CPU build: signature -> C driver.  One OpenMP thread block
# loop per kernel; the kernel bodies themselves are the reference's strings.
_CPU_DRIVERS = r"""
#define ORACLE_LAUNCH(call)                                              \
    _Pragma("omp parallel for schedule(static)")                          \
    for (long b = 0; b < grid; ++b) {                                    \
        blockIdx.x = (unsigned)b; blockDim.x = (unsigned)block;          \
        for (int t = 0; t < block; ++t) { threadIdx.x = (unsigned)t; call; } \
    }

extern "C" void ref_cpu_ppf_voting(long grid, int block,
        const float* points, const float* outputs, const float* probs, const int* point_idxs,
        float* grid_obj, const float* corner, float res, int n_ppfs, int n_rots,
        int gx, int gy, int gz, int adaptive) {
    ORACLE_LAUNCH(ppf_voting(points, outputs, probs, point_idxs, grid_obj, corner, res,
                             n_ppfs, n_rots, gx, gy, gz, adaptive != 0))
}
extern "C" void ref_cpu_backvote(long grid, int block,
        const float* points, const float* outputs, float* out_offsets, const int* point_idxs,
        const float* corner, float res, int n_ppfs, int n_rots, int gx, int gy, int gz,
        const float* gt_center, float tol) {
    ORACLE_LAUNCH(backvote(points, outputs, (float3*)out_offsets, point_idxs, corner, res,
                           n_ppfs, n_rots, gx, gy, gz, gt_center, tol))
}
extern "C" void ref_cpu_rot_voting(long grid, int block,
        const float* points, const float* not_used, const float* preds_rot, float* outputs_up,
        const int* point_idxs, const float* corner, float res, int n_ppfs, int n_rots,
        int gx, int gy, int gz) {
    ORACLE_LAUNCH(rot_voting(points, not_used, preds_rot, (float3*)outputs_up, point_idxs,
                             corner, res, n_ppfs, n_rots, gx, gy, gz))
}
extern "C" void ref_cpu_findpeak(long grid, int block,
        const float* grids, float* outputs, int width, int gx, int gy, int gz) {
    ORACLE_LAUNCH(findpeak(grids, outputs, width, gx, gy, gz))
}
"""


def extract_kernels(voting_py: str):
    """Return [(entry_name, cuda_source)] for every cp.RawKernel(...) in the file."""
    tree = ast.parse(open(voting_py).read())
    found = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "RawKernel":
            src = ast.literal_eval(node.args[0])
            name = ast.literal_eval(node.args[1])
            found.append((name, src))
    return found


def build(verbose: bool = False) -> bool:
    if not os.path.isdir(REF):
        if verbose:
            print(f"[oracle/_ref] {REF} absent -- keeping prebuilt files", file=sys.stderr)
        return False
    os.makedirs(OUT, exist_ok=True)
    kernels = extract_kernels(os.path.join(REF, "models", "voting.py"))
    assert sorted(n for n, _ in kernels) == ["backvote", "findpeak", "ppf_voting", "rot_voting"], kernels
    tmp = tempfile.mkdtemp(prefix="cppf_ref_")
    try:
        hdr = open(os.path.join(REF, "models", "include", "helper_math.cuh")).read()
        hdr = hdr.replace('#include "/usr/local/cuda-10.2/include/cuda_runtime.h"', "#include <cuda_runtime.h>")
        os.makedirs(os.path.join(tmp, "models", "src"))
        for dst in (os.path.join(tmp, "helper_math.cuh"), os.path.join(tmp, "models", "src", "helper_math.cuh")):
            with open(dst, "w") as f:
                f.write(hdr)
        # ---- GPU: one cubin per kernel, CuPy/NVRTC-like flags (no fast math, fmad on)
        for name, src in kernels:
            cu = os.path.join(tmp, f"{name}.cu")
            with open(cu, "w") as f:
                f.write(src)
            cmd = ["nvcc", "-arch=sm_100a", "-cubin", "-O3", "-w", "-I", tmp, "-o",
                   os.path.join(OUT, f"ref_{name}.cubin"), cu]
            subprocess.run(cmd, check=True, cwd=tmp)
        # ---- CPU: all four strings in one TU behind the shim
        cpp = os.path.join(tmp, "ref_cpu.cpp")
        with open(cpp, "w") as f:
            f.write(f'#include "{os.path.join(HERE, "ref_cpu_shim.h")}"\n')
            for name, src in kernels:
                f.write("#undef M_PI\n")
                f.write(src)
                f.write("\n")
            f.write(_CPU_DRIVERS)
        cmd = ["g++", "-O2", "-fopenmp", "-fPIC", "-shared", "-w", "-x", "c++",
               "-ffp-contract=fast", "-march=native",
               "-I", tmp, "-I", "/usr/local/cuda/include", cpp,
               "-o", os.path.join(OUT, "libref_voting_cpu.so")]
        subprocess.run(cmd, check=True, cwd=tmp)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if verbose:
        print("[oracle/_ref] built:", sorted(os.listdir(OUT)), file=sys.stderr)
    return True


if __name__ == "__main__":
    ok = build(verbose=True)
    sys.exit(0 if ok else 1)
