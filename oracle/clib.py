"""TEST INFRASTRUCTURE (oracle) -- numpy-facing loaders for the two CPU checkers.

* ``impl="oracle"``  -> ``oracle/_build/liboracle.so`` : the plain-C restatement in
  ``oracle/cppf_oracle.c`` (compiled here with gcc on first use / by ``build()``).
* ``impl="ref_cpu"`` -> ``oracle/_ref/libref_voting_cpu.so`` : the reference's own
  kernel strings compiled for the CPU by ``oracle/build_ref.py``.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may
import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(HERE, "_build")
_ORACLE_SO = os.path.join(_BUILD, "liboracle.so")
_REF_SO = os.path.join(HERE, "_ref", "libref_voting_cpu.so")
_libs = {}

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "cppf_oracle.c")
    if force or not os.path.exists(_ORACLE_SO) or os.path.getmtime(_ORACLE_SO) < os.path.getmtime(src):
        os.makedirs(_BUILD, exist_ok=True)
        subprocess.run(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-march=native", "-ffp-contract=fast",
                        src, "-lm", "-o", _ORACLE_SO], check=True)
    return _ORACLE_SO


def have_ref_cpu() -> bool:
    return os.path.exists(_REF_SO)


def _lib(impl):
    if impl in _libs:
        return _libs[impl]
    if impl == "oracle":
        L = C.CDLL(build())
        L.oracle_ppf_voting.argtypes = [_f32p, _f32p, _f32p, _i32p, _f32p, _f32p, C.c_float, C.c_long, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_ppf_voting_f64.argtypes = [_f32p, _f32p, _f32p, _i32p, _f64p, _f32p, C.c_float, C.c_long, C.c_int,
                                            C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_backvote.argtypes = [_f32p, _f32p, _f32p, _i32p, _f32p, C.c_float, C.c_long, C.c_int,
                                      C.c_int, C.c_int, C.c_int, _f32p, C.c_float]
        L.oracle_rot_voting.argtypes = [_f32p, _f32p, _f32p, _i32p, C.c_long, C.c_int]
        L.oracle_findpeak.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_grid_argmax.argtypes = [_f32p, C.c_long]
        L.oracle_grid_argmax.restype = C.c_long
        L.oracle_sphere_count.argtypes = [_f32p, C.c_long, _f32p, C.c_int, C.c_float, _i64p]
    elif impl == "ref_cpu":
        if not have_ref_cpu():
            raise FileNotFoundError(f"{_REF_SO} missing -- run `python oracle/build_ref.py` where /root/reference exists")
        L = C.CDLL(_REF_SO)
        L.ref_cpu_ppf_voting.argtypes = [C.c_long, C.c_int, _f32p, _f32p, _f32p, _i32p, _f32p, _f32p, C.c_float,
                                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_cpu_backvote.argtypes = [C.c_long, C.c_int, _f32p, _f32p, _f32p, _i32p, _f32p, C.c_float,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, C.c_float]
        L.ref_cpu_rot_voting.argtypes = [C.c_long, C.c_int, _f32p, _f32p, _f32p, _f32p, _i32p, _f32p, C.c_float,
                                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_cpu_findpeak.argtypes = [C.c_long, C.c_int, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int]
    else:
        raise ValueError(impl)
    _libs[impl] = L
    return L


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def ppf_voting(points, outputs, probs, idxs, grid_shape, corner, res, n_rots=72, adaptive=True,
               impl="oracle", f64=False, block=512):
    points, outputs, probs = _c(points, np.float32), _c(outputs, np.float32), _c(probs, np.float32)
    idxs, corner = _c(idxs, np.int32), _c(corner, np.float32)
    gx, gy, gz = (int(v) for v in grid_shape)
    P = idxs.shape[0]
    if f64:
        grid = np.zeros((gx, gy, gz), np.float64)
        _lib("oracle").oracle_ppf_voting_f64(points, outputs, probs, idxs, grid, corner, np.float32(res), P, n_rots,
                                             gx, gy, gz, int(bool(adaptive)))
        return grid
    grid = np.zeros((gx, gy, gz), np.float32)
    if impl == "oracle":
        _lib(impl).oracle_ppf_voting(points, outputs, probs, idxs, grid, corner, np.float32(res), P, n_rots,
                                     gx, gy, gz, int(bool(adaptive)))
    else:
        _lib(impl).ref_cpu_ppf_voting((P + block - 1) // block, block, points, outputs, probs, idxs, grid, corner,
                                      np.float32(res), P, n_rots, gx, gy, gz, int(bool(adaptive)))
    return grid


def backvote(points, outputs, idxs, grid_shape, corner, res, centre, tol, n_rots=72, impl="oracle", block=512):
    points, outputs = _c(points, np.float32), _c(outputs, np.float32)
    idxs, corner, centre = _c(idxs, np.int32), _c(corner, np.float32), _c(centre, np.float32)
    gx, gy, gz = (int(v) for v in grid_shape)
    P = idxs.shape[0]
    out = np.zeros((P, 3), np.float32)
    if impl == "oracle":
        _lib(impl).oracle_backvote(points, outputs, out, idxs, corner, np.float32(res), P, n_rots, gx, gy, gz,
                                   centre, np.float32(tol))
    else:
        _lib(impl).ref_cpu_backvote((P + block - 1) // block, block, points, outputs, out, idxs, corner,
                                    np.float32(res), P, n_rots, gx, gy, gz, centre, np.float32(tol))
    return out


def rot_voting(points, preds_rot, idxs, n_rots=72, impl="oracle", block=512):
    points, preds_rot, idxs = _c(points, np.float32), _c(preds_rot, np.float32), _c(idxs, np.int32)
    P = idxs.shape[0]
    out = np.zeros((P, n_rots, 3), np.float32)
    if impl == "oracle":
        _lib(impl).oracle_rot_voting(points, preds_rot, out, idxs, P, n_rots)
    else:
        dummy = np.zeros(3, np.float32)
        _lib(impl).ref_cpu_rot_voting((P + block - 1) // block, block, points, dummy, preds_rot, out, idxs, dummy,
                                      np.float32(0), P, n_rots, 0, 0, 0)
    return out


def findpeak(grid, width, literal=True, impl="oracle", block=512):
    g = _c(grid, np.float32)
    gx, gy, gz = g.shape
    out = np.zeros_like(g)
    if impl == "oracle":
        _lib(impl).oracle_findpeak(g, out, width, gx, gy, gz, int(bool(literal)))
    else:
        assert literal, "the reference string only has the literal (comma-operator) behaviour"
        _lib(impl).ref_cpu_findpeak((g.size + block - 1) // block, block, g, out, width, gx, gy, gz)
    return out


def grid_argmax(grid):
    g = _c(grid, np.float32)
    return int(_lib("oracle").oracle_grid_argmax(g.reshape(-1), g.size))


def sphere_count(cand, sphere, thr):
    cand, sphere = _c(cand, np.float32).reshape(-1, 3), _c(sphere, np.float32).reshape(-1, 3)
    counts = np.zeros(sphere.shape[0], np.int64)
    _lib("oracle").oracle_sphere_count(cand, cand.shape[0], sphere, sphere.shape[0], np.float32(thr), counts)
    return counts
