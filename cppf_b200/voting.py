"""Host-side mirror of the reference's ``models/voting.py``.

The reference exports four ``cupy.RawKernel`` objects and calls them as
``kernel((grid,1,1), (block,1,1), (args...))`` with the arguments in C-signature order
(``nocs/inference.py:197-205,221-228,268-275``).  The objects below have the same names
and call shape and run the sm_100a kernels of ``libcppf_b200.so``:

* array arguments may be torch CUDA tensors, anything exposing
  ``__cuda_array_interface__`` (CuPy), or numpy arrays (copied to the device; numpy
  *output* arguments are copied back, which synchronises);
* scalars may be Python or numpy/cupy scalars;
* the caller's grid/block hints are ignored (the reference over-launches N^2/512 blocks
  for P live threads, ``nocs/inference.py:192``).

The functional API underneath (``ppf_vote``, ``grid_argmax``, ``backvote``, ...) takes
torch tensors and is what ``cppf_b200.pipeline`` uses.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


class _CAIHolder:
    def __init__(self, obj):
        self.__cuda_array_interface__ = obj.__cuda_array_interface__


def _as_cuda(x, dtype, device=None):
    """-> (contiguous torch CUDA tensor of `dtype`, writeback) where writeback is the numpy
    array to copy results into (or None)."""
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("tensor arguments must live on a CUDA device (no CPU fallback)")
        t = x
    elif hasattr(x, "__cuda_array_interface__"):
        t = torch.as_tensor(_CAIHolder(x), device="cuda")
    elif isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x)).to(device or "cuda")
        out = t.to(dtype).contiguous()
        return out, x
    else:
        raise TypeError(f"unsupported array argument {type(x)}")
    if t.dtype != dtype or not t.is_contiguous():
        t = t.to(dtype).contiguous()
    return t, None


def _idx(x, device=None):
    """int32 or int64 [P,2] pair list -> (tensor, is64)."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x)).to(device or "cuda")
    elif not isinstance(x, torch.Tensor):
        x = torch.as_tensor(_CAIHolder(x), device="cuda")
    if x.dtype not in (torch.int32, torch.int64):
        x = x.to(torch.int64)
    x = x.contiguous()
    return x, int(x.dtype == torch.int64)


# ------------------------------------------------------------------------- functional API
def grid_dims(pc, corner, res):
    """Vote-grid dims of a CUDA cloud, nocs/inference.py:195: ``int((max - min) / res) + 1`` per axis with a TRUE float32
    division like NumPy's (and like geom_kernel on the device).  ``tensor / python_scalar`` on CUDA multiplies by the
    rounded reciprocal instead, which can differ by one ulp and flip the truncation; a device-tensor divisor does not."""
    ext = pc.max(0)[0] - corner
    return tuple(int(v) for v in (torch.div(ext, torch.full_like(ext, float(res))).int() + 1).cpu())


def ppf_vote(points, mu_nu, idxs, grid, corner, res, n_rots=72, adaptive=True, probs=None):
    """Centre voting (models/voting.py:8-66) accumulated into `grid` [gx,gy,gz] in place.
    idxs=None enumerates all N^2 ordered pairs."""
    dev = points.device
    n = points.shape[0]
    if idxs is None:
        ip, is64, n_pairs = None, 0, n * n
    else:
        idxs, is64 = _idx(idxs, dev)
        ip, n_pairs = idxs.data_ptr(), idxs.shape[0]
    assert mu_nu.shape[0] == n_pairs and mu_nu.is_contiguous() and grid.is_contiguous()
    gx, gy, gz = grid.shape
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cppf_ppf_vote(points.data_ptr(), mu_nu.data_ptr(), probs.data_ptr() if probs is not None else None,
                                            ip, is64, grid.data_ptr(), corner.data_ptr(), float(res), n, n_pairs,
                                            int(n_rots), gx, gy, gz, int(bool(adaptive)), _stream_ptr(dev)), "cppf_ppf_vote")
    return grid


def grid_argmax(grid, with_value=False):
    """First-max flat index in C order (np.argmax semantics, nocs/inference.py:208) -> int64[1] device tensor."""
    dev = grid.device
    out = torch.empty(1, dtype=torch.int64, device=dev)
    val = torch.empty(1, dtype=torch.float32, device=dev) if with_value else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cppf_grid_argmax(grid.data_ptr(), grid.numel(), out.data_ptr(),
                                               val.data_ptr() if with_value else None, _stream_ptr(dev)), "cppf_grid_argmax")
    return (out, val) if with_value else out


def backvote(points, mu_nu, idxs, grid_shape, corner, res, centre, tol, n_rots=72, want_offsets=True, want_mask=True):
    """Back-vote filter (models/voting.py:74-112) -> (offsets [P,3] | None, mask uint8 [P] | None)."""
    dev = points.device
    n = points.shape[0]
    if idxs is None:
        ip, is64, n_pairs = None, 0, n * n
    else:
        idxs, is64 = _idx(idxs, dev)
        ip, n_pairs = idxs.data_ptr(), idxs.shape[0]
    off = torch.zeros((n_pairs, 3), dtype=torch.float32, device=dev) if want_offsets else None
    mask = torch.empty(n_pairs, dtype=torch.uint8, device=dev) if want_mask else None
    gx, gy, gz = (int(v) for v in grid_shape)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cppf_backvote(points.data_ptr(), mu_nu.data_ptr(), off.data_ptr() if want_offsets else None,
                                            mask.data_ptr() if want_mask else None, ip, is64, corner.data_ptr(), float(res),
                                            n, n_pairs, int(n_rots), gx, gy, gz, centre.data_ptr(), float(tol),
                                            _stream_ptr(dev)), "cppf_backvote")
    return off, mask


def compact_pairs(mask, idxs, n_points, want_pos=False, want_idx=True):
    """point_idxs[mask] (nocs/inference.py:230-231), order preserving, on device.
    -> (idx int32 [P,2] (first `count` rows valid), count int64[1] device, pos | None)."""
    dev = mask.device
    n_pairs = mask.shape[0]
    if idxs is None:
        ip, is64 = None, 0
    else:
        idxs, is64 = _idx(idxs, dev)
        ip = idxs.data_ptr()
    L = _lib.lib()
    out = torch.empty((n_pairs, 2), dtype=torch.int32, device=dev) if want_idx else None
    pos = torch.empty(n_pairs, dtype=torch.int64, device=dev) if want_pos else None
    cnt = torch.empty(1, dtype=torch.int64, device=dev)
    scratch = torch.empty(int(L.cppf_compact_scratch_bytes(n_pairs)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.cppf_compact_pairs(mask.data_ptr(), ip, is64, int(n_points), n_pairs,
                                        out.data_ptr() if want_idx else None,
                                        pos.data_ptr() if want_pos else None, cnt.data_ptr(), scratch.data_ptr(),
                                        _stream_ptr(dev)), "cppf_compact_pairs")
    return out, cnt, pos


def rot_vote(points, preds_rot, idxs, n_rots=72, out=None):
    """Orientation candidates (models/voting.py:119-147) -> [P, n_rots, 3]."""
    dev = points.device
    idxs, is64 = _idx(idxs, dev)
    p = idxs.shape[0]
    if out is None:
        out = torch.zeros((p, n_rots, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cppf_rot_vote(points.data_ptr(), preds_rot.data_ptr(), out.data_ptr(), idxs.data_ptr(), is64,
                                            p, int(n_rots), _stream_ptr(dev)), "cppf_rot_vote")
    return out


def sphere_count(cand, sphere, thr, counts=None):
    """counts[s] += #{c : cand[c].sphere[s] > thr} (nocs/inference.py:282-283), int32 [n_bins]."""
    dev = cand.device
    cand = cand.reshape(-1, 3)
    if counts is None:
        counts = torch.zeros(sphere.shape[0], dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cppf_sphere_count(cand.data_ptr(), cand.shape[0], sphere.data_ptr(), sphere.shape[0],
                                                float(thr), counts.data_ptr(), _stream_ptr(dev)), "cppf_sphere_count")
    return counts


def findpeak(grid, width, literal=True):
    """models/voting.py:154-171 (literal=True reproduces the shipped comma-operator behaviour)."""
    dev = grid.device
    out = torch.empty_like(grid)
    gx, gy, gz = grid.shape
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cppf_findpeak(grid.data_ptr(), out.data_ptr(), int(width), gx, gy, gz, int(bool(literal)),
                                            _stream_ptr(dev)), "cppf_findpeak")
    return out


def sample_bins(logits, col0, n_bins, *, q=None, u=None, seed=None, stream_id=0, div=1.0, mul_a=1.0, mul_b=1.0, sub=0.0,
                want_bins=False):
    """softmax + categorical draw + decode (nocs/inference.py:185-188,245-256) over
    logits[:, col0:col0+n_bins].  Exactly one of q (Exp(1) noise [P,n_bins] -> torch.multinomial
    race), u (uniforms [P]) or seed (Philox) selects the sampler.  -> values f32 [P] (, bins)."""
    dev = logits.device
    assert logits.dim() == 2 and logits.is_contiguous()
    n = logits.shape[0]
    val = torch.empty(n, dtype=torch.float32, device=dev)
    bins = torch.empty(n, dtype=torch.int32, device=dev) if want_bins else None
    if q is not None:
        mode, noise = 0, q.contiguous()
        assert noise.shape == (n, n_bins)
    elif u is not None:
        mode, noise = 1, u.contiguous()
    else:
        assert seed is not None
        mode, noise = 2, None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cppf_sample_bins(logits.data_ptr(), n, logits.shape[1], int(col0), int(n_bins), mode,
                                               noise.data_ptr() if noise is not None else None, int(seed or 0),
                                               int(stream_id), float(div), float(mul_a), float(mul_b), float(sub),
                                               val.data_ptr(), 1, bins.data_ptr() if want_bins else None,
                                               _stream_ptr(dev)), "cppf_sample_bins")
    return (val, bins) if want_bins else val


# ------------------------------------------------------------------------- RawKernel-shaped objects
class _RawKernelShim:
    """Call shape of cupy.RawKernel: kernel(grid, block, args)."""
    name = ""

    def __call__(self, grid, block, args, **kwargs):
        self._run(*args)

    def __repr__(self):
        return f"<cppf_b200 kernel {self.name} (sm_100a, libcppf_b200.so)>"


def _writeback(pairs):
    for t, host in pairs:
        if host is not None:
            host[...] = t.cpu().numpy().reshape(host.shape)


class _PpfVoting(_RawKernelShim):
    name = "ppf_voting"

    def _run(self, points, outputs, probs, point_idxs, grid_obj, corner, res, n_ppfs, n_rots, gx, gy, gz, adaptive):
        pts, _ = _as_cuda(points, torch.float32)
        dev = pts.device
        out, _ = _as_cuda(outputs, torch.float32, dev)
        pr, _ = _as_cuda(probs, torch.float32, dev)
        idx, _ = _idx(point_idxs, dev)
        grid, gh = _as_cuda(grid_obj, torch.float32, dev)
        cor, _ = _as_cuda(corner, torch.float32, dev)
        n_ppfs = int(n_ppfs)
        g3 = grid.view(int(gx), int(gy), int(gz))
        ppf_vote(pts.view(-1, 3), out.view(-1, 2)[:n_ppfs], idx.view(-1, 2)[:n_ppfs], g3, cor, float(res), int(n_rots),
                 bool(adaptive), pr)
        _writeback([(grid, gh)])


class _Backvote(_RawKernelShim):
    name = "backvote"

    def _run(self, points, outputs, out_offsets, point_idxs, corner, res, n_ppfs, n_rots, gx, gy, gz, gt_center, tol):
        pts, _ = _as_cuda(points, torch.float32)
        dev = pts.device
        out, _ = _as_cuda(outputs, torch.float32, dev)
        off, oh = _as_cuda(out_offsets, torch.float32, dev)
        idx, is64 = _idx(point_idxs, dev)
        cor, _ = _as_cuda(corner, torch.float32, dev)
        ctr, _ = _as_cuda(gt_center, torch.float32, dev)
        n_ppfs = int(n_ppfs)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cppf_backvote(pts.data_ptr(), out.data_ptr(), off.data_ptr(), None, idx.data_ptr(), is64,
                                                cor.data_ptr(), float(res), pts.numel() // 3, n_ppfs, int(n_rots),
                                                int(gx), int(gy), int(gz), ctr.data_ptr(), float(tol), _stream_ptr(dev)),
                       "cppf_backvote")
        _writeback([(off, oh)])


class _RotVoting(_RawKernelShim):
    name = "rot_voting"

    def _run(self, points, not_used, preds_rot, outputs_up, point_idxs, corner, res, n_ppfs, n_rots, gx, gy, gz):
        pts, _ = _as_cuda(points, torch.float32)
        dev = pts.device
        rot, _ = _as_cuda(preds_rot, torch.float32, dev)
        up, uh = _as_cuda(outputs_up, torch.float32, dev)
        idx, _ = _idx(point_idxs, dev)
        n_ppfs = int(n_ppfs)
        rot_vote(pts.view(-1, 3), rot.view(-1)[:n_ppfs], idx.view(-1, 2)[:n_ppfs], int(n_rots),
                 out=up.view(-1, int(n_rots), 3)[:n_ppfs])
        _writeback([(up, uh)])


class _FindPeak(_RawKernelShim):
    name = "findpeak"
    literal = True      # behaviour of the string as shipped (models/voting.py:165-166)

    def _run(self, grids, outputs, width, gx, gy, gz):
        g, _ = _as_cuda(grids, torch.float32)
        dev = g.device
        o, oh = _as_cuda(outputs, torch.float32, dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cppf_findpeak(g.data_ptr(), o.data_ptr(), int(width), int(gx), int(gy), int(gz),
                                                int(self.literal), _stream_ptr(dev)), "cppf_findpeak")
        _writeback([(o, oh)])


ppf_kernel = _PpfVoting()
backvote_kernel = _Backvote()
rot_voting_kernel = _RotVoting()
findpeak_kernel = _FindPeak()
