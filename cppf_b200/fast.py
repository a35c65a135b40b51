"""Fused per-object path: thin torch-tensor wrappers over the `cppf_*` fast entry points
(include/cppf_b200.h, section "fused per-object path").  Bins, tail logits, the vote grid
and the survivor list stay on the device; nothing here synchronises."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .model import ROT_BINS, TR_BINS

HEAD_TR, HEAD_UP, HEAD_RIGHT, HEAD_TAIL = 1, 2, 4, 8


def _sp(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _idx_args(idxs):
    if idxs is None:
        return None, 0
    assert idxs.dtype in (torch.int32, torch.int64) and idxs.is_contiguous()
    return idxs.data_ptr(), int(idxs.dtype == torch.int64)


def decode_lut(vote_range, tr_bins=TR_BINS, rot_bins=ROT_BINS) -> torch.Tensor:
    """[136] fp32: value of every bin, built with the reference's own fp32 expression order
    (nocs/inference.py:187-188: b/(B-1)*2*vr0 - vr0, b/(B-1)*vr1; :252,256: b/(R-1)*pi)."""
    assert tr_bins == TR_BINS and rot_bins == ROT_BINS
    t = torch.arange(tr_bins).float()
    mu = t / (tr_bins - 1) * 2 * vote_range[0] - vote_range[0]
    nu = t / (tr_bins - 1) * vote_range[1]
    r = torch.arange(rot_bins).float() / (rot_bins - 1) * np.pi
    return torch.cat([mu, nu, r, r]).float().contiguous()


def encode_sample(ppf_encoder, pc, nrm, table, idxs, *, heads, uniforms=None, seed=0, bins=None, tail=None, impl="tc",
                  dbg_t=None, rows=None):
    """-> (bins uint8 [P,4], tail f32 [5,P] | None).  impl "tc": tcgen05 tensor-core encoder (3xTF32);
    "simt": fp32 FFMA warp-tile encoder.  rows=(lo, hi) with idxs=None: the dense pairs of rows [lo, hi) only (tc)."""
    dev = pc.device
    n = pc.shape[0]
    ip, is64 = _idx_args(idxs)
    row0 = 0
    if rows is not None:
        assert idxs is None and impl == "tc"
        row0 = int(rows[0])
    n_pairs = (n * n if rows is None else (int(rows[1]) - row0) * n) if idxs is None else idxs.shape[0]
    if bins is None:
        bins = torch.empty((n_pairs, 4), dtype=torch.uint8, device=dev)
    if (heads & HEAD_TAIL) and tail is None:
        tail = torch.empty((5, n_pairs), dtype=torch.float32, device=dev)
    if uniforms is not None:
        assert uniforms.shape == (n_pairs, 4) and uniforms.is_contiguous() and uniforms.dtype == torch.float32
    if impl == "tc":
        assert table.dim() == 3 and table.shape[0] * 4 == _lib.lib().cppf_tc_table_cols() and table.shape[1] == n, \
            "impl='tc' needs ppf_encoder.tc_preproject(feat)"
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cppf_encode_sample_tc_rows(
                pc.data_ptr(), nrm.data_ptr(), table.data_ptr(), ppf_encoder.tc_blob(dev).data_ptr(), ip, is64, n, n_pairs, row0,
                uniforms.data_ptr() if uniforms is not None else None, int(seed), int(heads), bins.data_ptr(),
                tail.data_ptr() if tail is not None else None, dbg_t.data_ptr() if dbg_t is not None else None,
                _sp(dev)), "cppf_encode_sample_tc")
        return bins, tail
    assert impl == "simt", impl
    assert table.dim() == 2 and table.shape[1] == 128, "impl='simt' needs ppf_encoder.preproject(feat)"
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cppf_encode_sample(
            pc.data_ptr(), nrm.data_ptr(), table.data_ptr(), ppf_encoder.weight_blob(dev).data_ptr(),
            ppf_encoder.head_blob(dev).data_ptr(), ip, is64, n, n_pairs,
            uniforms.data_ptr() if uniforms is not None else None, int(seed), int(heads), bins.data_ptr(),
            tail.data_ptr() if tail is not None else None, _sp(dev)), "cppf_encode_sample")
    return bins, tail


def vote_fits_private(dims) -> bool:
    return int(dims[0]) * int(dims[1]) * int(dims[2]) <= _lib.lib().cppf_vote_private_max_cells()


def vote_fast(points, idxs, grid, corner, res, *, mu_nu=None, bins=None, lut=None, n_rots=72, adaptive=True, scratch=None,
              rows=None):
    dev = points.device
    n = points.shape[0]
    ip, is64 = _idx_args(idxs)
    n_pairs = n * n if idxs is None else idxs.shape[0]
    gx, gy, gz = grid.shape
    if scratch is None:
        scratch = torch.empty(gx * gy * gz, dtype=torch.int64, device=dev)
    if rows is not None:                    # dense pairs of rows [lo, hi) (one object over several GPUs)
        assert idxs is None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cppf_vote_fast_rows(
                points.data_ptr(), mu_nu.data_ptr() if mu_nu is not None else None,
                bins.data_ptr() if bins is not None else None, lut.data_ptr() if lut is not None else None, int(rows[0]),
                grid.data_ptr(), scratch.data_ptr(), corner.data_ptr(), float(res), n, (int(rows[1]) - int(rows[0])) * n,
                int(n_rots), gx, gy, gz, int(bool(adaptive)), _sp(dev)), "cppf_vote_fast_rows")
        return grid
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cppf_vote_fast(
            points.data_ptr(), mu_nu.data_ptr() if mu_nu is not None else None,
            bins.data_ptr() if bins is not None else None, lut.data_ptr() if lut is not None else None, ip, is64,
            grid.data_ptr(), scratch.data_ptr(), corner.data_ptr(), float(res), n, n_pairs, int(n_rots), gx, gy, gz,
            int(bool(adaptive)), _sp(dev)), "cppf_vote_fast")
    return grid


def vote_slabs(points, idxs, grid, corner, res, *, mu_nu=None, bins=None, lut=None, n_rots=72, adaptive=True, scratch=None):
    """Centre vote for large grids by slab passes (cppf_vote_slabs): same kernel as vote_fast, one x-slab per CTA."""
    dev = points.device
    n = points.shape[0]
    ip, is64 = _idx_args(idxs)
    n_pairs = n * n if idxs is None else idxs.shape[0]
    gx, gy, gz = grid.shape
    if scratch is None:
        scratch = torch.empty(gx * gy * gz, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cppf_vote_slabs(
            points.data_ptr(), mu_nu.data_ptr() if mu_nu is not None else None,
            bins.data_ptr() if bins is not None else None, lut.data_ptr() if lut is not None else None, ip, is64,
            grid.data_ptr(), scratch.data_ptr(), corner.data_ptr(), float(res), n, n_pairs, int(n_rots), gx, gy, gz,
            int(bool(adaptive)), _sp(dev)), "cppf_vote_slabs")
    return grid


def vote_routed_supported(dims) -> bool:
    return bool(_lib.lib().cppf_vote_routed_supported(int(dims[0]), int(dims[1]), int(dims[2])))


def vote_routed(points, idxs, grid, corner, res, *, mu_nu=None, bins=None, lut=None, n_rots=72, adaptive=True, scratch=None):
    """Centre vote for grids of up to eight shared-memory slabs (cppf_vote_routed)."""
    dev = points.device
    n = points.shape[0]
    ip, is64 = _idx_args(idxs)
    n_pairs = n * n if idxs is None else idxs.shape[0]
    gx, gy, gz = grid.shape
    L = _lib.lib()
    if scratch is None:
        scratch = torch.empty(L.cppf_vote_routed_scratch_bytes(n_pairs, int(n_rots), gx, gy, gz), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.cppf_vote_routed(
            points.data_ptr(), mu_nu.data_ptr() if mu_nu is not None else None,
            bins.data_ptr() if bins is not None else None, lut.data_ptr() if lut is not None else None, ip, is64,
            grid.data_ptr(), scratch.data_ptr(), scratch.numel(), corner.data_ptr(), float(res), n, n_pairs, int(n_rots),
            gx, gy, gz, int(bool(adaptive)), _sp(dev)), "cppf_vote_routed")
    return grid


def backvote_bins(points, bins, lut, idxs, dims, corner, argmax_flat, res, tol, n_rots=72, mask=None, rows=None):
    dev = points.device
    n = points.shape[0]
    ip, is64 = _idx_args(idxs)
    n_pairs = (n * n if rows is None else (int(rows[1]) - int(rows[0])) * n) if idxs is None else idxs.shape[0]
    if mask is None:
        mask = torch.empty(n_pairs, dtype=torch.uint8, device=dev)
    if rows is not None:
        assert idxs is None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cppf_backvote_bins_rows(
                points.data_ptr(), bins.data_ptr(), lut.data_ptr(), int(rows[0]), mask.data_ptr(), corner.data_ptr(),
                argmax_flat.data_ptr(), float(res), float(tol), float(res), n, n_pairs, int(n_rots), int(dims[0]),
                int(dims[1]), int(dims[2]), _sp(dev)), "cppf_backvote_bins_rows")
        return mask
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cppf_backvote_bins(
            points.data_ptr(), bins.data_ptr(), lut.data_ptr(), ip, is64, mask.data_ptr(), corner.data_ptr(),
            argmax_flat.data_ptr(), float(res), float(tol), float(res), n, n_pairs, int(n_rots), int(dims[0]), int(dims[1]),
            int(dims[2]), _sp(dev)), "cppf_backvote_bins")
    return mask


def rot_hist(points, bins, lut, idxs, pos, count, sphere, *, which, n_rots=72, max_samples=10000, offset_seed=0, thr,
             counts=None, row0=None):
    dev = points.device
    ip, is64 = _idx_args(idxs)
    if counts is None:
        counts = torch.zeros(sphere.shape[0], dtype=torch.float32, device=dev)
    max_samples = max(1, min(int(max_samples), pos.shape[0]))
    if row0 is not None:
        assert idxs is None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cppf_rot_hist_rows(
                points.data_ptr(), bins.data_ptr(), lut.data_ptr(), int(row0), pos.data_ptr(), count.data_ptr(),
                sphere.data_ptr(), counts.data_ptr(), points.shape[0], int(n_rots), sphere.shape[0], int(which),
                int(max_samples), int(offset_seed), float(thr), _sp(dev)), "cppf_rot_hist_rows")
        return counts
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cppf_rot_hist(
            points.data_ptr(), bins.data_ptr(), lut.data_ptr(), ip, is64, pos.data_ptr(), count.data_ptr(),
            sphere.data_ptr(), counts.data_ptr(), points.shape[0], int(n_rots), sphere.shape[0], int(which),
            int(max_samples), int(offset_seed), float(thr), _sp(dev)), "cppf_rot_hist")
    return counts


def survivor_stats(points, nrm, tail, idxs, pos, count, sphere, best_up, best_right=None, row0=None):
    """-> float64 [6]: sum log-scale x3, survivor count, S_up, S_right."""
    dev = points.device
    ip, is64 = _idx_args(idxs)
    out = torch.empty(6, dtype=torch.float64, device=dev)
    if row0 is not None:
        assert idxs is None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cppf_survivor_stats_rows(
                points.data_ptr(), nrm.data_ptr(), tail.data_ptr(), int(row0), pos.data_ptr(), count.data_ptr(),
                sphere.data_ptr(), best_up.data_ptr(), best_right.data_ptr() if best_right is not None else None,
                out.data_ptr(), points.shape[0], tail.shape[1], _sp(dev)), "cppf_survivor_stats_rows")
        return out
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cppf_survivor_stats(
            points.data_ptr(), nrm.data_ptr(), tail.data_ptr(), ip, is64, pos.data_ptr(), count.data_ptr(),
            sphere.data_ptr(), best_up.data_ptr(), best_right.data_ptr() if best_right is not None else None,
            out.data_ptr(), points.shape[0], tail.shape[1], _sp(dev)), "cppf_survivor_stats")
    return out
