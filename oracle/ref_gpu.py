"""TEST INFRASTRUCTURE (oracle) -- launch the reference's OWN voting kernels on the GPU.

``oracle/build_ref.py`` compiles the CUDA-C strings of ``/root/reference/models/voting.py``
for sm_100a into ``oracle/_ref/ref_<name>.cubin``.  This module loads those cubins with
the CUDA driver API (cuda-python) and launches them on torch's current stream with the
reference's own launch shapes (``nocs/inference.py:192-205,216-228,267-275``), standing in
for the CuPy RawKernel objects (CuPy is not in the image).  Used by the ``-m gpu`` parity
tests and by bench.py's reference-GPU column only.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(HERE, "_ref")
_funcs = {}


def available() -> bool:
    try:
        from cuda.bindings import driver  # noqa: F401
    except Exception:
        return False
    return all(os.path.exists(os.path.join(_REF, f"ref_{n}.cubin")) for n in ("ppf_voting", "backvote", "rot_voting", "findpeak"))


def _check(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"CUDA driver error {err}")
    return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)


def _func(name):
    from cuda.bindings import driver
    if name not in _funcs:
        torch.cuda.init()
        torch.zeros(1, device="cuda")          # make sure the primary context is current
        data = open(os.path.join(_REF, f"ref_{name}.cubin"), "rb").read()
        mod = _check(driver.cuModuleLoadData(data))
        _funcs[name] = (_check(driver.cuModuleGetFunction(mod, name.encode())), mod)
    return _funcs[name][0]


def _launch(name, grid, block, args):
    """args: list of (ctype, value); device pointers as (c_void_p, int)."""
    from cuda.bindings import driver
    holders = [t(v) for t, v in args]
    ptrs = (ctypes.c_void_p * len(holders))(*[ctypes.addressof(h) for h in holders])
    stream = torch.cuda.current_stream().cuda_stream
    _check(driver.cuLaunchKernel(_func(name), grid, 1, 1, block, 1, 1, 0, stream, ctypes.addressof(ptrs), 0))


_P, _I, _F, _B = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_bool


def ppf_voting(points, outputs, probs, idxs32, grid, corner, res, n_rots, adaptive, over_launch=False):
    n_ppfs = idxs32.shape[0]
    n = points.shape[0]
    blocks = (n * n + 511) // 512 if over_launch else (n_ppfs + 511) // 512      # nocs/inference.py:192
    gx, gy, gz = grid.shape
    _launch("ppf_voting", blocks, 512, [(_P, points.data_ptr()), (_P, outputs.data_ptr()), (_P, probs.data_ptr()),
                                        (_P, idxs32.data_ptr()), (_P, grid.data_ptr()), (_P, corner.data_ptr()),
                                        (_F, res), (_I, n_ppfs), (_I, n_rots), (_I, gx), (_I, gy), (_I, gz),
                                        (_B, bool(adaptive))])
    return grid


def backvote(points, outputs, out_offsets, idxs32, corner, res, n_rots, grid_shape, centre, tol, n_threads=512):
    n_ppfs = idxs32.shape[0]
    gx, gy, gz = grid_shape
    _launch("backvote", (n_ppfs + n_threads - 1) // n_threads, n_threads,
            [(_P, points.data_ptr()), (_P, outputs.data_ptr()), (_P, out_offsets.data_ptr()), (_P, idxs32.data_ptr()),
             (_P, corner.data_ptr()), (_F, res), (_I, n_ppfs), (_I, n_rots), (_I, gx), (_I, gy), (_I, gz),
             (_P, centre.data_ptr()), (_F, tol)])
    return out_offsets


def rot_voting(points, preds_rot, outputs_up, idxs32, n_rots):
    n_ppfs = idxs32.shape[0]
    dummy = torch.zeros(4, device=points.device)
    _launch("rot_voting", (n_ppfs + 511) // 512, 512,
            [(_P, points.data_ptr()), (_P, dummy.data_ptr()), (_P, preds_rot.data_ptr()), (_P, outputs_up.data_ptr()),
             (_P, idxs32.data_ptr()), (_P, dummy.data_ptr()), (_F, 0.0), (_I, n_ppfs), (_I, n_rots), (_I, 0), (_I, 0), (_I, 0)])
    return outputs_up


def findpeak(grid, out, width):
    gx, gy, gz = grid.shape
    n = grid.numel()
    _launch("findpeak", (n + 511) // 512, 512, [(_P, grid.data_ptr()), (_P, out.data_ptr()), (_I, width), (_I, gx), (_I, gy), (_I, gz)])
    return out
