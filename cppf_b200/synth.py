"""Seeded synthetic inputs for parity tests and benchmarks.

There are no datasets or checkpoints in the image, so every cloud, pair list, weight
set and vote target is generated here from explicit seeds.  Pure numpy; no GPU.
"""
from __future__ import annotations

import numpy as np

# config/category/bottle.yaml:2-8, config/config.yaml:7-8,21 of the reference
BOTTLE = dict(category="bottle", res=4e-3, vote_range=(0.25, 0.25), scale_mean=(0.05, 0.15, 0.05),
              up_sym=True, regress_right=False, z_right=False, tr_num_bins=32, rot_num_bins=36, knn=60)
# config/category/chair.yaml (SUN RGB-D scale)
CHAIR = dict(category="chair", res=3e-2, vote_range=(0.7863312261283193, 0.7863312261283193),
             scale_mean=(0.29636475108453486, 0.4208450279047945, 0.2789450356430997),
             up_sym=False, regress_right=True, z_right=False, tr_num_bins=32, rot_num_bins=36, knn=60)


def synth_bottle(n: int, seed: int = 0, scale: float = 1.0):
    """Bottle-like cloud, y-up, centred so that the object origin is the vote target.
    80 % body (r=0.035, y in [-0.11,0.05]), 20 % neck (r=0.015, y in [0.05,0.11]);
    outward radial normals.  Returns (pc[n,3] f32, normals[n,3] f32)."""
    rng = np.random.default_rng(seed)
    nb = int(round(0.8 * n))
    th = rng.uniform(0.0, 2 * np.pi, n)
    y = np.concatenate([rng.uniform(-0.11, 0.05, nb), rng.uniform(0.05, 0.11, n - nb)])
    r = np.concatenate([np.full(nb, 0.035), np.full(n - nb, 0.015)])
    pc = np.stack([r * np.cos(th), y, r * np.sin(th)], -1) * scale
    nrm = np.stack([np.cos(th), np.zeros(n), np.sin(th)], -1)
    return pc.astype(np.float32), nrm.astype(np.float32)


def synth_cylinder_grid64(n: int, seed: int = 0, res: float = 4e-3, cells: int = 64):
    """Cylinder with diameter = height = (cells-0.5)*res including the six axis-extreme
    points, so that int(extent/res)+1 == cells on every axis (the 64^3 vote-grid config)."""
    rng = np.random.default_rng(seed)
    half = 0.5 * (cells - 0.5) * res
    th = rng.uniform(0.0, 2 * np.pi, n)
    y = rng.uniform(-half, half, n)
    th[:4] = [0.0, np.pi, 0.5 * np.pi, 1.5 * np.pi]
    y[4], y[5] = -half, half
    pc = np.stack([half * np.cos(th), y, half * np.sin(th)], -1)
    nrm = np.stack([np.cos(th), np.zeros(n), np.sin(th)], -1)
    return pc.astype(np.float32), nrm.astype(np.float32)


def sample_pairs(n: int, p: int, seed: int = 0) -> np.ndarray:
    """Random ordered pairs like nocs/inference.py:177 but seeded. int64 [p,2]."""
    return np.random.default_rng(seed + 1).integers(0, n, (p, 2)).astype(np.int64)


def dense_pairs(n: int) -> np.ndarray:
    """All ordered pairs, row-major (i = point a, j = point b), including i == j."""
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    return np.stack([i, j], -1).reshape(-1, 2).astype(np.int64)


def vote_grid_geometry(pc: np.ndarray, res: float):
    """nocs/inference.py:194-195: corner = min, grid = int((max-min)/res)+1 per axis."""
    t = np.ascontiguousarray(np.asarray(pc).T)           # [3, N]: numpy reduces a [N, 3] array along axis 0 ~25x slower
    lo, hi = t.min(1), t.max(1)
    dims = ((hi - lo) / res).astype(np.int32) + 1
    return lo.astype(np.float32), tuple(int(d) for d in dims)


def trained_like_tr(pc: np.ndarray, idxs: np.ndarray, tr_num_bins: int = 32, vote_range=(0.25, 0.25)):
    """(mu, nu) a well-trained network would emit: the geometric targets of
    utils/dataset.py:27-36 snapped to the bin centres of nocs/inference.py:187-188."""
    a = pc[idxs[:, 0]].astype(np.float64)
    d = a - pc[idxs[:, 1]].astype(np.float64)
    du = d / (np.linalg.norm(d, axis=-1, keepdims=True) + 1e-7)
    mu = np.sum(a * du, -1)
    nu = np.linalg.norm(a - mu[:, None] * du, axis=-1)
    b_mu = np.clip(np.rint((mu + vote_range[0]) / (2 * vote_range[0]) * (tr_num_bins - 1)), 0, tr_num_bins - 1)
    b_nu = np.clip(np.rint(nu / vote_range[1] * (tr_num_bins - 1)), 0, tr_num_bins - 1)
    mu_q = (b_mu.astype(np.float32) / np.float32(tr_num_bins - 1) * np.float32(2 * vote_range[0])
            - np.float32(vote_range[0]))
    nu_q = b_nu.astype(np.float32) / np.float32(tr_num_bins - 1) * np.float32(vote_range[1])
    return np.stack([mu_q, nu_q], -1).astype(np.float32)


def trained_like_rot(pc: np.ndarray, idxs: np.ndarray, rot_num_bins: int = 36, up_sym: bool = True):
    """Up-axis angle target (utils/dataset.py:38-45) snapped to the rot bins (inference.py:252)."""
    d = pc[idxs[:, 0]].astype(np.float64) - pc[idxs[:, 1]].astype(np.float64)
    du = d / (np.linalg.norm(d, axis=-1, keepdims=True) + 1e-7)
    ang = np.arccos(np.clip(du[:, 1], -1, 1))
    if up_sym:
        ang = np.minimum(ang, np.pi - ang)
    b = np.clip(np.rint(ang / np.pi * (rot_num_bins - 1)), 0, rot_num_bins - 1)
    return (b.astype(np.float32) / np.float32(rot_num_bins - 1) * np.float32(np.pi)).astype(np.float32)


def trained_like_bins_dense_torch(pc, cfg, chunk_rows: int = 256):
    """Device-side version of trained_like_tr / trained_like_rot for ALL ordered pairs of a cloud
    (row-major): uint8 [N*N, 3] = (mu bin, nu bin, up bin).  Used to give the vote stage the load
    a trained network would produce (SURVEY.md section 8d i) when only random-init weights exist."""
    import torch
    n = pc.shape[0]
    tb, rb = cfg["tr_num_bins"], cfg["rot_num_bins"]
    vr0, vr1 = cfg["vote_range"]
    out = torch.empty((n * n, 3), dtype=torch.uint8, device=pc.device)
    pcd = pc.double()
    for r0 in range(0, n, chunk_rows):
        a = pcd[r0:r0 + chunk_rows, None, :]
        d = a - pcd[None, :, :]
        du = d / (d.norm(dim=-1, keepdim=True) + 1e-7)
        mu = (a * du).sum(-1)
        nu = (a - mu[..., None] * du).norm(dim=-1)
        ang = torch.arccos(du[..., 1].clamp(-1, 1))
        if cfg["up_sym"]:
            ang = torch.minimum(ang, np.pi - ang)
        b = torch.stack([((mu + vr0) / (2 * vr0) * (tb - 1)).round().clamp(0, tb - 1),
                         (nu / vr1 * (tb - 1)).round().clamp(0, tb - 1),
                         (ang / np.pi * (rb - 1)).round().clamp(0, rb - 1)], -1)
        out[r0 * n:(r0 + a.shape[0]) * n] = b.reshape(-1, 3).to(torch.uint8)
    return out
