"""ctypes binding of ``libcppf_b200.so`` (C ABI in ``include/cppf_b200.h``).

There is deliberately no fallback: if the shared library is missing or a call returns a
CUDA error, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CPPF_B200_LIB: developer override used by tools/build_variants.py to A/B compile-time variants of the library
LIB_PATH = os.environ.get("CPPF_B200_LIB") or os.path.join(_HERE, "libcppf_b200.so")
_lib = None

_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_f = C.c_float

# name -> (restype, argtypes); mirrors include/cppf_b200.h line by line
SIGNATURES = {
    "cppf_abi_version": (_i, []),
    "cppf_error_string": (C.c_char_p, [_i]),
    "cppf_launch_count": (C.c_uint64, []),
    "cppf_ppf_blob_floats": (_i, [_i]),
    "cppf_ppf_feat_dim": (_i, []),
    "cppf_ppf_preproject": (_i, [_p, _p, _p, _i, _p]),
    "cppf_ppf_encode": (_i, [_p, _p, _p, _p, _p, _i, _p, _p, _i, _i64, _i, _i, _i, _p]),
    "cppf_sample_bins": (_i, [_p, _i64, _i, _i, _i, _i, _p, C.c_uint64, C.c_uint32, _f, _f, _f, _f, _p, _i, _p, _p]),
    "cppf_ppf_vote": (_i, [_p, _p, _p, _p, _i, _p, _p, _f, _i, _i64, _i, _i, _i, _i, _i, _p]),
    "cppf_grid_argmax": (_i, [_p, _i64, _p, _p, _p]),
    "cppf_backvote": (_i, [_p, _p, _p, _p, _p, _i, _p, _f, _i, _i64, _i, _i, _i, _i, _p, _f, _p]),
    "cppf_compact_scratch_bytes": (_i64, [_i64]),
    "cppf_compact_pairs": (_i, [_p, _p, _i, _i, _i64, _p, _p, _p, _p, _p]),
    "cppf_compact_count": (_i, [_p, _i64, _p, _p, _p]),
    "cppf_rot_hist_mask": (_i, [_p, _p, _p, _p, _i, _p, _p, _i64, _p, _p, _p, _i, _i, _i, _i, _i64, C.c_uint64, _f, _p, _p]),
    "cppf_survivor_stats_mask": (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _i, _i64, _p]),
    "cppf_rot_vote": (_i, [_p, _p, _p, _p, _i, _i64, _i, _p]),
    "cppf_sphere_count": (_i, [_p, _i64, _p, _i, _f, _p, _p]),
    "cppf_findpeak": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "cppf_head_blob_floats": (_i, []),
    "cppf_encode_sample": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i64, _p, C.c_uint64, _i, _p, _p, _p]),
    "cppf_tc_blob_floats": (_i, []),
    "cppf_tc_table_cols": (_i, []),
    "cppf_tc_preproject": (_i, [_p, _p, _p, _i, _p]),
    "cppf_encode_sample_tc": (_i, [_p, _p, _p, _p, _p, _i, _i, _i64, _p, C.c_uint64, _i, _p, _p, _p, _p]),
    "cppf_pe_blob_floats": (_i, []),
    "cppf_knn": (_i, [_p, _i, _i, _p, _p]),
    "cppf_point_encode": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "cppf_vote_scratch_bytes": (_i64, [_i, _i, _i]),
    "cppf_vote_private_max_cells": (_i, []),
    "cppf_vote_fast": (_i, [_p, _p, _p, _p, _p, _i, _p, _p, _p, _f, _i, _i64, _i, _i, _i, _i, _i, _p]),
    "cppf_vote_finalize": (_i, [_p, _p, _i64, _p]),
    "cppf_vote_routed_supported": (_i, [_i, _i, _i]),
    "cppf_vote_routed_scratch_bytes": (_i64, [_i64, _i, _i, _i, _i]),
    "cppf_vote_routed": (_i, [_p, _p, _p, _p, _p, _i, _p, _p, _i64, _p, _f, _i, _i64, _i, _i, _i, _i, _i, _p]),
    "cppf_vote_slabs_supported": (_i, [_i, _i, _i]),
    "cppf_vote_slabs": (_i, [_p, _p, _p, _p, _p, _i, _p, _p, _p, _f, _i, _i64, _i, _i, _i, _i, _i, _p]),
    "cppf_backvote_bins": (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _f, _f, C.c_double, _i, _i64, _i, _i, _i, _i, _p]),
    "cppf_rot_hist": (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i64, C.c_uint64, _f, _p]),
    "cppf_survivor_stats": (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _i, _i64, _p]),
    "cppf_backproject_scratch_bytes": (_i64, [_i, _i]),
    "cppf_backproject": (_i, [_p, _i, _p, _i, _i, _p, C.c_double, _p, _p, _p, _p, _p]),
    "cppf_voxel_scratch_bytes": (_i64, [_i64]),
    "cppf_voxel_first": (_i, [_p, _p, _i64, C.c_double, _p, _p, _p, _p, _p]),
    "cppf_normals_pca": (_i, [_p, _i, _i, _i, _p, _p, _p]),
    "cppf_pair_filter": (_i, [_p, _p, _p, _i, _i, _i64, _p, _p]),
    "cppf_gaussian3d": (_i, [_p, _p, _p, _i, _i, _i, C.c_double, C.c_double, _p]),
    "cppf_scene_proposals": (_i, [_p, _i, _i, _i, _f, _i, _f, _i, _p, _p, _p]),
    "cppf_pose_record_doubles": (_i, []),
    "cppf_pose_args_bytes": (_i, []),
    "cppf_pose_workspace_bytes": (_i64, [_i, _i64, _i, _i, _i, _i, _i, _i64]),
    "cppf_pose_fused": (_i, [_p, _p]),
    "cppf_pose_batch": (_i, [_p, _i, _i, _i, _p]),
    "cppf_timing_create": (_p, []),
    "cppf_timing_destroy": (None, [_p]),
    "cppf_timing_reserve": (_i, [_p, _i]),
    "cppf_timing_stages": (_i, []),
    "cppf_timing_stage_name": (C.c_char_p, [_i]),
    "cppf_timing_collect": (_i, [_p, _p]),
    "cppf_encode_sample_tc_rows": (_i, [_p, _p, _p, _p, _p, _i, _i, _i64, _i, _p, C.c_uint64, _i, _p, _p, _p, _p]),
    "cppf_vote_fast_rows": (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _f, _i, _i64, _i, _i, _i, _i, _i, _p]),
    "cppf_backvote_bins_rows": (_i, [_p, _p, _p, _i, _p, _p, _p, _f, _f, C.c_double, _i, _i64, _i, _i, _i, _i, _p]),
    "cppf_rot_hist_rows": (_i, [_p, _p, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i64, C.c_uint64, _f, _p]),
    "cppf_survivor_stats_rows": (_i, [_p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _i, _i64, _p]),
    "cppf_vote_count": (_i, [_p, _p, _p, _p, _p, _i, _p, _f, _i, _i64, _i, _i, _i, _i, _i, _p, _p]),
    "cppf_peak_shared_atomics": (_i, [_i, _i, _i, _i, _i, _p, _p]),
    "cppf_peak_global_red": (_i, [_i, _i, _i, _i, _p, _p]),
}


class PoseArgs(C.Structure):
    """struct cppf_pose_args of include/cppf_b200.h, field for field."""
    _fields_ = [
        ("struct_bytes", _i64),
        ("pc", _p), ("nrm", _p), ("idx", _p), ("pe_blob", _p), ("tc_blob", _p), ("lut", _p), ("sphere", _p),
        ("uniforms", _p), ("inject_bins", _p), ("workspace", _p), ("record", _p), ("timing", _p),
        ("n_pairs", _i64), ("workspace_bytes", _i64), ("rot_subsample", _i64), ("seed", C.c_uint64), ("res_host", C.c_double),
        ("n_points", _i), ("idx_is_64", _i), ("knn", _i), ("n_rots", _i), ("adaptive", _i), ("regress_right", _i),
        ("n_sphere", _i), ("inject_cols", _i), ("max_cells", _i), ("routed_max_cells", _i), ("sample_pairs", _i),
        ("res", _f), ("tol", _f), ("cos_thr", _f),
    ]


def lib():
    """The loaded library; raises loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m cppf_b200.build` "
                "(cppf_b200 has no CPU or PyTorch fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code: int, what: str = "cppf_b200") -> None:
    if code != 0:
        msg = lib().cppf_error_string(code)
        raise RuntimeError(f"{what} failed: CUDA error {code} ({msg.decode() if msg else '?'})")


def launch_count() -> int:
    return int(lib().cppf_launch_count())
