"""TEST INFRASTRUCTURE (oracle) -- the reference's per-object script body, ``nocs/inference.py:174-339``, restated as ONE
CPU function over the other oracle pieces (``ref_model``: the torch modules restated; ``clib``: the voting kernels as plain C,
or the reference's own CUDA-C strings built for the CPU when ``oracle/_ref`` travelled).

Used by (i) ``tests/test_gpu_pipeline_parity.py`` -- ``PoseEstimator.estimate`` against this function on injected noise: centre
argmax exact, RT / scales within 1e-4 relative -- and (ii) ``bench.py``'s CPU arm, which only times it.  Never imported by
``cppf_b200/``.

The reference is unseeded (``np.random.randint`` :177, ``torch.multinomial`` :186/:246-255, ``np.random.shuffle`` :279); every
random choice here can be injected through ``noise`` so that two implementations can be compared draw for draw:

  q_mu, q_nu   float32 [P, tr_num_bins]   Exp(1) variates: ATen's multinomial(p, 1) IS argmax(p / q) with q ~ Exp(1)
  q_up, q_right float32 [P, rot_num_bins] same for the rotation heads, indexed by the ORIGINAL pair row (survivors take
                                          the rows they came from)
  sub_key      float32 [P]                the 10 000-survivor sub-sample of :277-281 = the survivors with the smallest
                                          keys (a uniformly random subset, like the shuffle)

Anything not injected is drawn from ``torch.Generator().manual_seed(seed)`` / ``np.random.default_rng(seed)``.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import clib, ref_model


def _race(logits, q, gen):
    """softmax + multinomial(., 1) (nocs/inference.py:185-186): the exponential race, injected or drawn."""
    if q is None:
        q = torch.empty(logits.shape, dtype=torch.float32).exponential_(1.0, generator=gen)
    return ref_model.sample_bins_race(logits, torch.as_tensor(q))


def exact_knn(pc: torch.Tensor, k: int) -> torch.Tensor:
    """Neighbour sets of models/model.py:47 from EXACT distances (float64), not torch.cdist's |x|^2+|y|^2-2x.y form whose
    rounding decides near-ties of the k-th neighbour differently from any other implementation (SURVEY.md 8a notes)."""
    d = pc.double()
    out = torch.empty((pc.shape[0], k), dtype=torch.long)
    for r0 in range(0, pc.shape[0], 512):
        dd = ((d[r0:r0 + 512, None, :] - d[None, :, :]) ** 2).sum(-1)
        out[r0:r0 + 512] = torch.topk(dd, k, largest=False, sorted=False)[1]
    return out


def estimate(pc, nrm, sd_pe, sd_ppf, idxs, cfg, noise=None, seed=0, sphere=None, impl=None, cdist_knn=False,
             return_debug=False, timings=None):
    """pc, nrm float32 [N,3]; sd_*: state_dicts of the reference modules; idxs int [P,2] (:177); cfg: dict with the category
    constants (config/category/*.yaml + config/config.yaml).  -> dict(RT, scales, T, flat, n_survivors, best_bins, ...).
    timings: optional dict that receives the wall-clock seconds of the per-OBJECT stages ("point_encoder": kNN + SPRIN,
    O(N k); "orientation": :258-284 on the 10 000-survivor sub-sample, a constant once more than 10 000 pairs survive) and
    of the per-PAIR stages ("pairs": everything else)."""
    t_start = time.perf_counter()
    noise = noise or {}
    impl = impl or ("ref_cpu" if clib.have_ref_cpu() else "oracle")
    gen = torch.Generator().manual_seed(int(seed))
    n = pc.shape[0]
    B, RB = cfg["tr_num_bins"], cfg["rot_num_bins"]
    tpc, tn = torch.from_numpy(np.ascontiguousarray(pc)), torch.from_numpy(np.ascontiguousarray(nrm))
    idxs = np.asarray(idxs)
    # ---- :180-182 point features, first encoder pass
    if cdist_knn:
        dist = torch.cdist(tpc[None], tpc[None])[0]                                   # :180 literally
        feat = ref_model.point_encode(tpc, tn, dist, sd_pe, cfg["knn"])
    else:
        feat = ref_model.point_encode_nbrs(tpc, tn, exact_knn(tpc, cfg["knn"]), sd_pe)
    t_pe = time.perf_counter()
    logits = ref_model.ppf_encode_idx(tpc, tn, feat, idxs, sd_ppf)
    # ---- :183-188 sample (mu, nu)
    b_mu = _race(logits[:, :B], noise.get("q_mu"), gen)
    b_nu = _race(logits[:, B:2 * B], noise.get("q_nu"), gen)
    tr = ref_model.decode_tr(b_mu, b_nu, B, cfg["vote_range"]).numpy()
    # ---- :191-211 centre vote, argmax
    lo, hi = np.min(pc, 0), np.max(pc, 0)
    dims = ((hi - lo) / cfg["res"]).astype(np.int32) + 1                              # :195
    idx32 = idxs.astype(np.int32)
    n_rots = int(cfg.get("num_rots", 72))
    grid = clib.ppf_voting(pc, tr, np.ones(n, np.float32), idx32, dims, lo, cfg["res"], n_rots,
                           bool(cfg.get("adaptive_voting", True)), impl=impl)
    flat, T_est = ref_model.centre_from_grid(grid, lo, cfg["res"])
    # ---- :216-231 back-vote filter
    oc = clib.backvote(pc, tr, idx32, dims, lo, cfg["res"], T_est.astype(np.float32), np.float32(3 * cfg["res"]), n_rots,
                       impl=impl)
    mask = np.any(oc != 0, -1)
    pos = np.nonzero(mask)[0]
    kept = idxs[mask]
    out = {"T": T_est, "flat": int(flat), "dims": tuple(int(v) for v in dims), "n_survivors": int(len(kept)), "impl": impl}
    if return_debug:
        out.update(grid=grid, tr=tr, mask=mask, feat=feat)
    if len(kept) == 0:
        if timings is not None:
            timings.update(point_encoder=t_pe - t_start, orientation=0.0, pairs=time.perf_counter() - t_pe,
                           orientation_is_fixed=False)
        return out
    # ---- :236-256 second encoder pass on the survivors
    l2 = ref_model.ppf_encode_idx(tpc, tn, feat, kept, sd_ppf)
    log_scale = l2[:, -3:].mean(0).numpy()                                            # :335
    sphere = ref_model.fibonacci_sphere(int(4 * np.pi / (cfg.get("angle_prec", 1.5) / 180 * np.pi))) if sphere is None else sphere
    thr = np.cos(cfg.get("angle_prec", 1.5) / 180 * np.pi)
    dirs, bests = [], []
    t_orient = 0.0
    for j, (c0, aux_col, tag) in enumerate([(2 * B, -5, "up"), (2 * B + RB, -4, "right")]):
        if j == 1 and not cfg.get("regress_right", False):                            # :260-261
            continue
        q = noise.get(f"q_{tag}")
        b_rot = _race(l2[:, c0:c0 + RB], None if q is None else torch.as_tensor(q)[pos], gen)
        rot = ref_model.decode_rot(b_rot, RB).numpy().astype(np.float32)              # :252,256
        m = int(cfg.get("rot_subsample", 10000) or 0)
        if m and len(kept) > m:                                                       # :277-281
            key = noise.get("sub_key")
            if key is None:
                sel = np.random.default_rng(seed + 17 + j).permutation(len(kept))[:m]
            else:
                sel = np.argsort(np.asarray(key)[pos], kind="stable")[:m]
        else:
            sel = np.arange(len(kept))
        t_o = time.perf_counter()
        cand = clib.rot_voting(pc, rot[sel], kept[sel].astype(np.int32), n_rots, impl=impl)        # :265-275
        counts = ((torch.from_numpy(cand.reshape(-1, 3)) @ torch.from_numpy(sphere.T.astype(np.float32))) >
                  float(np.float32(thr))).sum(0).numpy()                              # :282-283
        best = int(np.argmax(counts))                                                 # :284
        t_orient += time.perf_counter() - t_o
        final, _, _ = ref_model.aux_sign(pc, nrm, kept, sphere[best], l2[:, aux_col].numpy())      # :286-302
        dirs.append(final)
        bests.append(best)
        if return_debug:
            out[f"counts_{tag}"] = counts
    # ---- :305-339 pose
    RT, scales = ref_model.assemble_pose(dirs[0], dirs[1] if len(dirs) > 1 else None, T_est, log_scale, cfg["scale_mean"],
                                         z_right=cfg.get("z_right", False), regress_right=cfg.get("regress_right", False),
                                         scale_mul=cfg.get("scale_mul", 2.0))
    out.update(RT=RT, scales=scales, up=dirs[0], best_bins=bests, log_scale=log_scale)
    if timings is not None:
        timings.update(point_encoder=t_pe - t_start, orientation=t_orient, pairs=time.perf_counter() - t_pe - t_orient,
                       orientation_is_fixed=bool(int(cfg.get("rot_subsample", 10000) or 0) and
                                                 len(kept) > int(cfg.get("rot_subsample", 10000) or 0)))
    return out
