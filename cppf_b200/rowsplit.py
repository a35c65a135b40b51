"""One object over several GPUs ("optional second axis").

Objects normally shard whole (``shard.py``).  For ONE huge object (dense N >= 16 k: 2.7e8 ordered pairs)
the rows of the pair matrix are split instead: rank r takes the pairs (a, b) with a in its row block and
every b, all ranks hold the whole cloud (N x 24 B) and compute the per-point features redundantly (O(N k)),
and the path gets three small exchange steps, each a sum:

  1. the vote grid -- as the exact 64-bit fixed-point sums the vote kernels leave in their scratch
     (``cppf_vote_finalize`` converts after the reduction), so the grid, hence the argmax, is bit for bit the
     grid of a single-GPU run over the same bins;
  2. the orientation histogram(s) (480 integer-valued float32 counts per direction);
  3. the survivor statistics (6 float64 sums: log-scale x3, count, aux-sign scores).

Between the exchanges every rank runs the same kernels as the single-GPU staged path on its own pairs: in dense
ROW-BLOCK mode (the ``cppf_*_rows`` entry points enumerate rows [lo, hi) of the pair matrix in-kernel, keeping the
row-aligned MMA tiles and the tiled vote batches of the dense path and drawing the same Philox stream as a
full-matrix launch) when the grid fits the shared-memory vote kernel, else over an int32 pair list of the block.  The steps are written as a generator that yields each
tensor to be summed, so the same code runs under ``torch.distributed`` (``estimate_rowsplit``; NCCL on the
box, one process per GPU) and in a single process that plays all ranks in lockstep
(``estimate_rowsplit_local``: the parity test on one GPU).

The reference has no counterpart (one process, one object at a time: ``nocs/inference.py:120``); the
semantics are those of ``nocs/inference.py:177-339`` with ``point_idxs`` = all ordered pairs.  The random
10 000-pair sub-sample of the orientation vote (``:276``) is drawn per rank (ceil(10 000 / world) each).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, fast, voting


def row_block(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [lo, hi) of the N x N pair matrix owned by `rank`: contiguous, sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("rank outside the world")
    base, rem = divmod(int(n), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def block_pairs(n: int, lo: int, hi: int, device=None) -> torch.Tensor:
    """int32 [(hi - lo) * n, 2]: the pairs (a, b), a in [lo, hi), b in [0, n), row-major -- entry p of the block is
    entry lo * n + p of the dense enumeration (unit of work)."""
    a = torch.arange(lo, hi, dtype=torch.int32, device=device).repeat_interleave(n)
    b = torch.arange(n, dtype=torch.int32, device=device).repeat(hi - lo)
    return torch.stack([a, b], 1).contiguous()


_PAIRS_CACHE = {}


def _cached_block_pairs(n, lo, hi, dev):
    """The block's pair list only depends on (N, rows): keep the last one per device (1 GB at N = 16 k, world 2)."""
    key = (str(dev), n, lo, hi)
    hit = _PAIRS_CACHE.get(str(dev))
    if hit is None or hit[0] != key:
        _PAIRS_CACHE[str(dev)] = None
        hit = _PAIRS_CACHE[str(dev)] = (key, block_pairs(n, lo, hi, dev))
    return hit[1]


def _steps(est, pc, nrm, seed, uniforms, inject_bins, rank, world, out):
    """Generator over the per-rank stages; yields every tensor that has to be summed over the ranks (in place)."""
    cfg, dev = est.cfg, est.device
    n = pc.shape[0]
    corner = pc.min(0)[0]
    dims = voting.grid_dims(pc, corner, cfg.res)                                            # nocs/inference.py:194-195
    cells = dims[0] * dims[1] * dims[2]
    lo, hi = row_block(n, world, rank)
    sl = slice(lo * n, hi * n)
    feat = est.point_features(pc, nrm)
    table = est.ppf.tc_preproject(feat) if est.encoder_impl == "tc" else est.ppf.preproject(feat)
    heads = fast.HEAD_TR | fast.HEAD_UP | fast.HEAD_TAIL | (fast.HEAD_RIGHT if cfg.regress_right else 0)
    # Dense row-block mode (the `_rows` entry points: the kernels enumerate rows [lo, hi) of the pair matrix themselves, with
    # every dense-mode shortcut and no pair list in HBM) whenever the grid fits the shared-memory vote kernel and the
    # tcgen05 encoder runs; otherwise an explicit int32 pair list of the block (1 GB at N = 16 k, world 2).
    dense_rows = (est.encoder_impl == "tc" and cfg.num_rots <= 72 and fast.vote_fits_private(dims) and hi > lo)
    rows = (lo, hi) if dense_rows else None
    row0 = lo if dense_rows else None
    idxs = None if dense_rows else _cached_block_pairs(n, lo, hi, dev)
    # the Philox stream is keyed by the pair's index in the whole matrix (dense rows): every world size draws the same bins
    bins, tail = fast.encode_sample(est.ppf, pc, nrm, table, idxs, heads=heads,
                                    uniforms=uniforms[sl] if uniforms is not None else None,
                                    seed=int(seed) if dense_rows else int(seed) * 1000003 + rank, impl=est.encoder_impl, rows=rows)
    if inject_bins is not None:
        bins[:, :inject_bins.shape[1]] = inject_bins[sl]
    grid = torch.zeros(dims, dtype=torch.float32, device=dev)
    exact = cfg.num_rots <= 72 and (fast.vote_fits_private(dims) or fast.vote_routed_supported(dims))
    if exact:
        junk = torch.zeros(dims, dtype=torch.float32, device=dev)
        if fast.vote_fits_private(dims):
            acc = torch.zeros(cells, dtype=torch.int64, device=dev)
            fast.vote_fast(pc, idxs, junk, corner, cfg.res, bins=bins, lut=est.lut, n_rots=cfg.num_rots,
                           adaptive=cfg.adaptive_voting, scratch=acc, rows=rows)
        else:
            nb = _lib.lib().cppf_vote_routed_scratch_bytes(idxs.shape[0], cfg.num_rots, *dims)
            scratch = torch.zeros((nb + 7) // 8, dtype=torch.int64, device=dev)
            fast.vote_routed(pc, idxs, junk, corner, cfg.res, bins=bins, lut=est.lut, n_rots=cfg.num_rots,
                             adaptive=cfg.adaptive_voting, scratch=scratch.view(torch.uint8))
            acc = scratch[:cells]
        yield acc                                                   # exchange 1: exact integer vote sums
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cppf_vote_finalize(acc.data_ptr(), grid.data_ptr(), cells,
                                                     torch.cuda.current_stream(dev).cuda_stream), "cppf_vote_finalize")
    else:                                                           # scene-scale grids: fp32 partial grids
        b = bins.long()
        mu_nu = torch.stack([est.lut[b[:, 0]], est.lut[32 + b[:, 1]]], -1).contiguous()
        voting.ppf_vote(pc, mu_nu, idxs, grid, corner, cfg.res, cfg.num_rots, cfg.adaptive_voting)
        yield grid
    flat = voting.grid_argmax(grid)
    mask = fast.backvote_bins(pc, bins, est.lut, idxs, dims, corner, flat, cfg.res, 3 * cfg.res, cfg.num_rots, rows=rows)
    _, cnt, pos = voting.compact_pairs(mask, idxs, n, want_pos=True, want_idx=False)
    n_dirs = 2 if cfg.regress_right else 1
    quota = -(-int(cfg.rot_subsample) // world) if cfg.rot_subsample else (1 << 40)
    counts = torch.zeros((n_dirs, est.sphere.shape[0]), dtype=torch.float32, device=dev)
    for j in range(n_dirs):
        fast.rot_hist(pc, bins, est.lut, idxs, pos, cnt, est.sphere, which=j, n_rots=cfg.num_rots, max_samples=quota,
                      offset_seed=(seed * 7919 + j) * 64 + rank, thr=est.cos_thr, counts=counts[j], row0=row0)
    yield counts                                                    # exchange 2: orientation histograms
    bests = [voting.grid_argmax(counts[j]) for j in range(n_dirs)]
    stats = fast.survivor_stats(pc, nrm, tail, idxs, pos, cnt, est.sphere, bests[0], bests[1] if cfg.regress_right else None,
                                row0=row0)
    yield stats                                                     # exchange 3: survivor statistics
    rec = torch.cat([flat.double()] + [b.double() for b in bests] + [stats, corner.double()])
    out.update(record=rec, n_dirs=n_dirs, grid=grid, bins=bins, mask=mask, rows=(lo, hi), dims=dims)


def _prepare(est, pc_in, nrm_in):
    dev = est.device
    pc = torch.as_tensor(pc_in).to(dev, torch.float32).contiguous()
    nrm = torch.as_tensor(nrm_in).to(dev, torch.float32).contiguous()
    return pc, nrm


@torch.no_grad()
def estimate_rowsplit(est, pc_in, nrm_in, seed: int = 0, uniforms=None, inject_bins=None, group=None, return_debug=False):
    """Pose of one object whose pair rows are split over the ranks of `group` (every rank passes the same cloud and
    gets the same pose).  uniforms / inject_bins, if given, cover ALL N^2 pairs (row-major) on every rank."""
    if not dist.is_initialized():
        rank, world = 0, 1
    else:
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    pc, nrm = _prepare(est, pc_in, nrm_in)
    out = {}
    for t in _steps(est, pc, nrm, seed, uniforms, inject_bins, rank, world, out):
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    pose = est._pose_from_record(out["record"].cpu().numpy(), out["n_dirs"], out["dims"])
    if return_debug:
        pose.update(grid=out["grid"], bins=out["bins"], mask=out["mask"], rows=out["rows"])
    return pose


@torch.no_grad()
def estimate_rowsplit_local(est, pc_in, nrm_in, world: int, seed: int = 0, uniforms=None, inject_bins=None, return_debug=False):
    """The same computation with all `world` ranks played by this process on this GPU, in lockstep (each exchange
    is a plain sum of the ranks' tensors).  For tests and for sizing the split; no speed-up."""
    pc, nrm = _prepare(est, pc_in, nrm_in)
    outs = [{} for _ in range(world)]
    gens = [_steps(est, pc, nrm, seed, uniforms, inject_bins, r, world, outs[r]) for r in range(world)]
    while True:
        ts = []
        for g in gens:
            try:
                ts.append(next(g))
            except StopIteration:
                pass
        if not ts:
            break
        if len(ts) != world:
            raise RuntimeError("ranks fell out of step")
        total = ts[0].clone()
        for t in ts[1:]:
            total += t
        for t in ts:
            t.copy_(total)
    recs = [o["record"].cpu().numpy() for o in outs]
    for r in recs[1:]:
        if not np.array_equal(r, recs[0]):
            raise RuntimeError("ranks disagree on the pose record")
    pose = est._pose_from_record(recs[0], outs[0]["n_dirs"], outs[0]["dims"])
    if return_debug:
        pose.update(grid=outs[0]["grid"], bins=torch.cat([o["bins"] for o in outs]), mask=torch.cat([o["mask"] for o in outs]))
    return pose
