"""TEST INFRASTRUCTURE (oracle) -- NOT part of the shipped product path.

CPU restatement (torch fp32 tensor algebra, functional style, weights passed as a
plain ``state_dict``) of the PyTorch half of the reference hot path and of the host
glue of ``nocs/inference.py``.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / reference arm may import it.

Pinning: ``tests/test_oracle_golden.py`` checks every function here against
fixtures minted by *importing the reference's own modules*
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``) and, when ``/root/reference``
is present, against the live reference modules as well.

All citations are relative to ``/root/reference``.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------- pair MLP (a1-a4)
def _linear(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def res_layer(x, sd, prefix):
    """models/model.py:26-31 -- y = fc2(relu(fc1(x))) + (fc0(x) | x); no norm (bn off, :11)."""
    skip = _linear(x, sd, prefix + ".fc0") if (prefix + ".fc0.weight") in sd else x
    return _linear(torch.relu(_linear(x, sd, prefix + ".fc1")), sd, prefix + ".fc2") + skip


def n_res_layers(sd):
    return len({k.split(".")[1] for k in sd if k.startswith("res_layers.")})


def pair_mlp(x, sd):
    """models/model.py:134-137 -- the ResLayer stack then the `final` Linear."""
    for i in range(n_res_layers(sd)):
        x = res_layer(x, sd, f"res_layers.{i}")
    return _linear(x, sd, "final")


def ppf_encode_idx(pc, nrm, feat, idxs, sd):
    """models/model.py:117-137 forward_with_idx.  pc,nrm [N,3], feat [N,F], idxs [P,2] -> [P,out]."""
    idxs = idxs.long() if isinstance(idxs, torch.Tensor) else torch.as_tensor(np.asarray(idxs)).long().to(pc.device)
    ia, ib = idxs[:, 0], idxs[:, 1]
    d = pc[ia] - pc[ib]                                     # :120  a minus b
    dn = torch.norm(d, dim=-1)                              # :121
    dh = d / (dn[:, None] + 1e-7)                           # :122
    ppf = torch.stack([(nrm[ia] * dh).sum(-1), (nrm[ib] * dh).sum(-1),
                       (nrm[ia] * nrm[ib]).sum(-1), dn], -1)        # :124-129
    return pair_mlp(torch.cat([feat[ia], feat[ib], ppf], -1), sd)   # :132-137


def ppf_encode_dense(pc, nrm, feat, dist, sd):
    """models/model.py:92-115 dense branch for one object.  dist [N,N] comes from the
    caller (torch.cdist at nocs/inference.py:180).  -> [N,N,out], row i = point a."""
    n = pc.shape[0]
    d = pc[:, None, :] - pc[None, :, :]                     # :92  xx[i,j] = pc[i]-pc[j]
    dh = d / (dist[..., None] + 1e-7)                       # :93
    ppf = torch.stack([(nrm[:, None, :] * dh).sum(-1), (nrm[None, :, :] * dh).sum(-1),
                       (nrm[:, None, :] * nrm[None, :, :]).sum(-1), dist], -1)   # :100-105
    x = torch.cat([feat[:, None, :].expand(n, n, -1), feat[None, :, :].expand(n, n, -1), ppf], -1)  # :107
    return pair_mlp(x, sd)                                  # :110-115 (row chunking is memory-only)


# ---------------------------------------------------------------- point encoder (a5)
def rifeat(nbr_pts, centre):
    """models/sprin.py:40-60.  nbr_pts [N,K,3] absolute coords, centre [N,1,3] -> [N,K,6]."""
    mean = nbr_pts.mean(-2, keepdim=True)
    l1, l2, l3 = mean - nbr_pts, nbr_pts - centre, centre - mean
    n1 = torch.norm(l1, dim=-1, keepdim=True)
    n2 = torch.norm(l2, dim=-1, keepdim=True)
    n3 = torch.norm(l3, dim=-1, keepdim=True).expand_as(n2)
    t1 = (l1 * l2).sum(-1, keepdim=True) / (n1 * n2 + 1e-7)
    t2 = (l2 * l3).sum(-1, keepdim=True) / (n2 * n3 + 1e-7)
    t3 = (l3 * l1).sum(-1, keepdim=True) / (n3 * n1 + 1e-7)
    return torch.cat([n1, n2, n3, t1, t2, t3], -1)


def sprin_kernel_mlp(x, sd, prefix):
    """models/sprin.py:63-71 -- Linear, LayerNorm(eps 1e-5), ReLU per hidden unit; last Linear bare.
    Sequential indices: Linear at 0,3,6,..; LayerNorm at 1,4,7,..."""
    lin_ids = sorted({int(k[len(prefix) + 1:].split(".")[0]) for k in sd
                      if k.startswith(prefix + ".") and k.endswith(".weight") and sd[k].dim() == 2})
    for j, li in enumerate(lin_ids):
        x = _linear(x, sd, f"{prefix}.{li}")
        if j + 1 < len(lin_ids):
            w, b = sd[f"{prefix}.{li + 1}.weight"], sd[f"{prefix}.{li + 1}.bias"]
            x = torch.relu(F.layer_norm(x, (w.shape[0],), w, b, 1e-5))
    return x


def point_encode_nbrs(pc, nrm, nbrs_idx, sd):
    """models/model.py:63-77 forward_nbrs with num_layers=1 (nocs/inference.py:82):
    gather k neighbours, SparseSO3Conv (models/sprin.py:94-107), GlobalInfoProp (:80-83).
    pc,nrm [N,3]; nbrs_idx [N,K] long -> feat [N, out + out//4]."""
    nb = pc[nbrs_idx]                                                    # :64
    nb_norm = torch.norm(nb - pc[:, None, :], dim=-1, keepdim=True)      # :65-66
    nb_cos = (nrm[nbrs_idx] * nrm[:, None, :]).sum(-1, keepdim=True)     # :68-69
    nbr_feat = torch.cat([nb_norm, nb_cos], -1)                          # :71
    kern = sprin_kernel_mlp(rifeat(nb, pc[:, None, :]), sd, "spconvs.0.kernel")     # sprin.py:96
    contracted = torch.einsum("nkr,nki->nri", kern, nbr_feat).flatten(-2)           # sprin.py:98
    conv = _linear(contracted, sd, "spconvs.0.outnet")                              # sprin.py:99
    w, b = sd["spconvs.0.layer_norm.weight"], sd["spconvs.0.layer_norm.bias"]
    conv = F.layer_norm(conv, (w.shape[0],), w, b, 1e-5)                            # sprin.py:105
    tran = _linear(conv, sd, "aggrs.0.linear")                                      # sprin.py:80
    glob = tran.max(0, keepdim=True)[0].expand(conv.shape[0], -1)                   # sprin.py:82
    return torch.cat([conv, glob], -1)


def knn_from_dist(dist, k):
    """models/model.py:47 -- k smallest per row, self included, order unspecified."""
    return torch.topk(dist, k, largest=False, sorted=False)[1]


def point_encode(pc, nrm, dist, sd, k):
    """models/model.py:46-61 forward."""
    return point_encode_nbrs(pc, nrm, knn_from_dist(dist, k), sd)


# ---------------------------------------------------------------- decode / sampling glue (a6)
def sample_bins_race(logits, q):
    """nocs/inference.py:185-186.  softmax then torch.multinomial(.,1); ATen's multinomial
    is argmax(p / q) with q ~ Exp(1) drawn from the generator, so with q
    injected the draw is reproducible.  logits,q [P,bins] -> long [P]."""
    return torch.argmax(torch.softmax(logits, -1) / q, -1)


def sample_bins_cdf(logits, u, exp2=False):
    """Inverse-CDF categorical draw: same distribution as torch.multinomial, one uniform
    per row.  This is the product kernels' native sampler (cppf_b200/csrc), restated:
    e_k = exp(l_k - max l); pick the first k with cumsum(e)[k] > u * sum(e), computed in
    fp32 sequentially in bin order; clamp to the last bin.  exp2=True restates the fused
    kernel's form e_k = 2^((l_k - max) * log2 e)."""
    l = logits.float()
    if exp2:
        e = torch.exp2((l - l.max(-1, keepdim=True)[0]) * 1.4426950408889634)
    else:
        e = torch.exp(l - l.max(-1, keepdim=True)[0])
    tot = torch.zeros(l.shape[0], dtype=torch.float32)
    for k in range(l.shape[1]):
        tot = tot + e[:, k]
    t = u.float() * tot
    acc = torch.zeros_like(tot)
    out = torch.full((l.shape[0],), l.shape[1] - 1, dtype=torch.long)
    done = torch.zeros(l.shape[0], dtype=torch.bool)
    for k in range(l.shape[1]):
        acc = acc + e[:, k]
        hit = (acc > t) & ~done
        out[hit] = k
        done |= hit
    return out


def decode_tr(bin_mu, bin_nu, tr_num_bins, vote_range):
    """nocs/inference.py:187-188.  fp32: mu = b/(B-1)*2*vr0 - vr0 ; nu = b/(B-1)*vr1."""
    mu = bin_mu.float() / (tr_num_bins - 1) * 2 * vote_range[0] - vote_range[0]
    nu = bin_nu.float() / (tr_num_bins - 1) * vote_range[1]
    return torch.stack([mu, nu], -1)


def decode_rot(bin_rot, rot_num_bins):
    """nocs/inference.py:252,256.  angle = b/(B-1)*pi (fp32 tensor times python float)."""
    return bin_rot.float() / (rot_num_bins - 1) * np.pi


# ---------------------------------------------------------------- sphere bins / targets (a13, a14)
def fibonacci_sphere(samples):
    """utils/util.py:102-118, float64 python arithmetic."""
    phi = math.pi * (3.0 - math.sqrt(5.0))
    pts = []
    for i in range(samples):
        y = 1 - (i / float(samples - 1)) * 2
        r = math.sqrt(1 - y * y)
        th = phi * i
        pts.append((math.cos(th) * r, y, math.sin(th) * r))
    return np.array(pts)


def generate_target_tr(pc, point_idxs):
    """utils/dataset.py:27-36 -- ground-truth (mu, nu) of an object centred at the origin."""
    a, b = pc[point_idxs[:, 0]].astype(np.float64), pc[point_idxs[:, 1]].astype(np.float64)
    d = a - b
    du = d / (np.linalg.norm(d, axis=-1, keepdims=True) + 1e-7)
    mu = np.sum(a * du, -1)
    nu = np.linalg.norm(a - mu[:, None] * du, axis=-1)
    return np.stack([mu, nu], -1).astype(np.float32)


def generate_target_rot(pc, point_idxs, up_sym):
    """utils/dataset.py:38-51 (up axis only): angle between the pair direction and +y."""
    a, b = pc[point_idxs[:, 0]].astype(np.float64), pc[point_idxs[:, 1]].astype(np.float64)
    d = a - b
    du = d / (np.linalg.norm(d, axis=-1, keepdims=True) + 1e-7)
    ang = np.arccos(np.clip(du[:, 1], -1, 1))
    if up_sym:
        ang = np.minimum(ang, np.arccos(np.clip(-du[:, 1], -1, 1)))
    return ang.astype(np.float32)


# ---------------------------------------------------------------- host pose glue (a8, a11)
def centre_from_grid(grid, corner, res):
    """nocs/inference.py:207-211.  first-max argmax in C order -> world coordinates (float64)."""
    flat = int(np.argmax(grid, axis=None))
    cell = np.array(np.unravel_index(flat, grid.shape))
    return flat, np.asarray(corner, dtype=np.float64) + cell * res


def aux_sign(pc, nrm, point_idxs, best_dir, aux_logits):
    """nocs/inference.py:286-302.  Decide the sign of best_dir from the aux head by the lower
    BCE-with-logits against target = [flipped normal . best_dir > 0]."""
    ab = pc[point_idxs[:, 0]] - pc[point_idxs[:, 1]]
    abn = ab / (np.sqrt(np.sum(ab ** 2, -1)) + 1e-7)[..., None]
    pn = nrm[point_idxs[:, 0]].copy()
    pn[np.sum(pn * abn, -1) < 0] *= -1
    target = torch.from_numpy((np.sum(pn * best_dir, -1) > 0).astype(np.float32))
    aux = torch.as_tensor(aux_logits).float()
    up = F.binary_cross_entropy_with_logits(aux, target).item()
    down = F.binary_cross_entropy_with_logits(aux, 1.0 - target).item()
    return (-best_dir if down < up else best_dir), up, down


def assemble_pose(up, right, T, log_scale_mean, scale_mean, z_right=False, regress_right=False, scale_mul=2.0):
    """nocs/inference.py:305-339 (laptop fix-up :314-323 excluded: out of scope)."""
    up = np.asarray(up, dtype=np.float64)
    if regress_right:
        right = np.asarray(right, dtype=np.float64).copy()
        right -= np.dot(up, right) * up
        right /= (np.linalg.norm(right) + 1e-9)
    else:
        right = np.array([0, -up[2], up[1]])
        right /= (np.linalg.norm(right) + 1e-9)
    if z_right:
        R = np.stack([np.cross(up, right), up, right], -1)
    else:
        R = np.stack([right, up, np.cross(right, up)], -1)
    pred_scale = np.exp(np.asarray(log_scale_mean, dtype=np.float32)) * np.asarray(scale_mean) * scale_mul
    sn = np.linalg.norm(pred_scale)
    RT = np.eye(4, dtype=np.float32)
    RT[:3, :3] = R * sn
    RT[:3, 3] = T
    return RT, (pred_scale / sn).astype(np.float32)
