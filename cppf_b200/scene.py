"""Scene-scale ("zero-shot") mode: the reference's ``nocs/zero_shot.ipynb`` over the
cppf_b200 kernels -- a whole depth frame, ~5 M random pairs, a regression-head pair network (``out_dim = 9``: mu, nu,
up angle, -, up-aux, -, log-scale x3), one scene-sized vote grid, Gaussian smoothing, greedy multi-peak proposals, then a
per-proposal back-vote / instance segmentation / refinement.  "cell N" = N-th code cell of the notebook."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, voting
from .pipeline import fibonacci_sphere


def _sp(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def pair_filter(pc, nrm, idxs) -> torch.Tensor:
    """cell 6 -> uint8 [P] (1 = keep)."""
    dev = pc.device
    idxs = idxs.contiguous()
    keep = torch.empty(idxs.shape[0], dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cppf_pair_filter(pc.data_ptr(), nrm.data_ptr(), idxs.data_ptr(), int(idxs.dtype == torch.int64),
                                               pc.shape[0], idxs.shape[0], keep.data_ptr(), _sp(dev)), "cppf_pair_filter")
    return keep


def gaussian_filter(grid: torch.Tensor, sigma: float = 1.0, truncate: float = 4.0) -> torch.Tensor:
    """scipy.ndimage.gaussian_filter(grid, sigma) (cell 9) for a float32 CUDA grid [gx,gy,gz]."""
    assert grid.dtype == torch.float32 and grid.dim() == 3 and grid.is_contiguous()
    out, tmp = torch.empty_like(grid), torch.empty_like(grid)
    with torch.cuda.device(grid.device):
        _lib.check(_lib.lib().cppf_gaussian3d(grid.data_ptr(), out.data_ptr(), tmp.data_ptr(), *grid.shape, float(sigma),
                                              float(truncate), _sp(grid.device)), "cppf_gaussian3d")
    return out


def scene_proposals(smoothed: torch.Tensor, thresh: float = 50.0, margin: int = 10, rel_stop: float = 0.7, max_props: int = 64):
    """cell 9 greedy proposals; `smoothed` is modified in place like the notebook's array.
    -> list of (loc int[3], value, contrast)."""
    assert smoothed.dtype == torch.float32 and smoothed.dim() == 3 and smoothed.is_contiguous()
    host = (C.c_float * (5 * max_props))()
    scratch = torch.zeros(64, dtype=torch.uint8, device=smoothed.device)
    with torch.cuda.device(smoothed.device):
        n = _lib.lib().cppf_scene_proposals(smoothed.data_ptr(), *smoothed.shape, float(thresh), int(margin), float(rel_stop),
                                            int(max_props), host, scratch.data_ptr(), _sp(smoothed.device))
    if n < 0:
        _lib.check(-n, "cppf_scene_proposals")
    return [(np.array([int(host[5 * i]), int(host[5 * i + 1]), int(host[5 * i + 2])]), float(host[5 * i + 3]), float(host[5 * i + 4]))
            for i in range(n)]


@torch.no_grad()
def estimate_scene(point_encoder, ppf_encoder, pc, nrm, high_res_pc=None, high_res_nrm=None, subset=None, *, res: float,
                   scale_mean, knn: int = 60, n_pairs: int = 5_000_000, num_rots: int = 72, angle_tol: float = 2.0,
                   thresh: float = 50.0, margin: int = 10, min_contrib: int = 12, rot_subsample: int = 10000, seed: int = 0,
                   idxs=None):
    """nocs/zero_shot.ipynb cells 5-11.  pc, nrm: the sparse scene cloud (float32 CUDA [N,3]); optionally the
    high-resolution cloud it was sub-sampled from (`subset` = indices, cell 3) on which the point features are computed
    (cell 7).  ppf_encoder: regression head, out_dim 9.  -> list of dicts (T, up, RT, scales, n_pairs, instance mask)."""
    dev = pc.device
    n = pc.shape[0]
    g = torch.Generator(device=dev).manual_seed(seed)
    if idxs is None:
        idxs = torch.randint(0, n, (n_pairs, 2), generator=g, device=dev, dtype=torch.int32)        # cell 5
    keep = pair_filter(pc, nrm, idxs)                                                              # cell 6
    idxs = idxs[keep.bool()].contiguous()
    if high_res_pc is not None:                                                                    # cell 7
        feat = point_encoder.encode_fused(high_res_pc, high_res_nrm)[subset]
    else:
        feat = point_encoder.encode_fused(pc, nrm)
    preds = ppf_encoder.forward_with_idx(pc, nrm, feat, idxs)                                      # [P,9], unbatched like models/model.py:117-137
    preds_tr = preds[:, :2].contiguous()
    corner = pc.min(0)[0]                                                                          # cell 8
    dims = voting.grid_dims(pc, corner, res)
    grid = torch.zeros(dims, dtype=torch.float32, device=dev)
    voting.ppf_vote(pc, preds_tr, idxs, grid, corner, res, num_rots, True)
    smoothed = gaussian_filter(grid, 1.0)                                                          # cell 9
    props = scene_proposals(smoothed, thresh, margin)
    n_bins = int(4 * np.pi / (angle_tol / 180 * np.pi))                                            # cell 1
    sphere_np = fibonacci_sphere(n_bins)
    sphere = torch.from_numpy(sphere_np.astype(np.float32)).to(dev)
    cos_thr = float(np.float32(np.cos(angle_tol / 180 * np.pi)))
    out = []
    for loc, cnt, diff in props:                                                                   # cell 11
        T_est = corner.double().cpu().numpy() + loc * res
        centre = torch.from_numpy(T_est.astype(np.float32)).to(dev)
        _, mask = voting.backvote(pc, preds_tr, idxs, dims, corner, res, centre, 3 * res, num_rots, want_offsets=False)
        sel = idxs[mask.bool()]
        contrib = torch.bincount(sel.reshape(-1).long(), minlength=n)                              # instance segmentation
        keep_pt = contrib > min_contrib
        keep_pair = keep_pt[sel[:, 0].long()] | keep_pt[sel[:, 1].long()]
        sel = sel[keep_pair].contiguous()
        if sel.shape[0] == 0:
            continue
        pm = ppf_encoder.forward_with_idx(pc, nrm, feat, sel)
        rot = pm[:, 2].contiguous()
        sub = sel
        if sel.shape[0] > rot_subsample:
            pick = torch.randperm(sel.shape[0], generator=g, device=dev)[:rot_subsample]
            sub, rot = sel[pick].contiguous(), rot[pick].contiguous()
        cand = voting.rot_vote(pc, rot, sub, num_rots)
        counts = voting.sphere_count(cand, sphere, cos_thr)
        best_up = sphere_np[int(torch.argmax(counts).item())]
        a_i, b_i = sel[:, 0].long(), sel[:, 1].long()                                              # aux classification
        ab = pc[a_i] - pc[b_i]
        abn = ab / (ab.pow(2).sum(-1).sqrt() + 1e-7)[:, None]
        pn = nrm[a_i]
        pn = torch.where(((pn * abn).sum(-1) < 0)[:, None], -pn, pn)
        target = ((pn * torch.from_numpy(best_up.astype(np.float32)).to(dev)).sum(-1) > 0).float()
        bce = torch.nn.functional.binary_cross_entropy_with_logits
        up = -best_up if bce(pm[:, 4], 1.0 - target) < bce(pm[:, 4], target) else best_up
        right = np.array([0, -up[2], up[1]])
        right /= np.linalg.norm(right)
        R = np.stack([right, up, np.cross(right, up)], -1)
        scale3 = (torch.exp(pm[:, -3:]) * torch.tensor(scale_mean, dtype=torch.float32, device=dev) * 2).mean(0).cpu().numpy()
        sn = np.linalg.norm(scale3)
        RT = np.eye(4)
        RT[:3, :3] = R * sn
        RT[:3, 3] = T_est
        out.append(dict(T=T_est, up=up, RT=RT, scales=scale3 / sn, votes=cnt, contrast=diff, n_pairs=int(sel.shape[0]),
                        instance_mask=keep_pt))
    return out
