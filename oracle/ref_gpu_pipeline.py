"""TEST / BENCH INFRASTRUCTURE (oracle) -- the reference's per-object flow ON THE GPU, timed stage by stage.

BASELINE.json config 2 is "fused PPF+MLP kernel vs reference CuPy"; SURVEY.md 2.1 sets the GPU bar as the reference's own
CUDA C recompiled for sm_100a.  This module is that arm: the reference's voting kernels exactly as they are in
``models/voting.py`` (the strings compiled to ``oracle/_ref/ref_*.cubin`` by ``oracle/build_ref.py``, launched with the
reference's launch shapes, ``nocs/inference.py:192-205,216-228,267-275``) + the reference's torch modules restated in
``oracle/ref_model.py`` (pinned to ``models/model.py`` / ``models/sprin.py`` by the golden tests) run by torch-CUDA -- the
same ATen / cuBLAS kernels the reference modules launch -- in the order and with the host round trips of
``nocs/inference.py:174-339``.  Only bench.py and the GPU tests import it; nothing under ``cppf_b200/`` does.

Pairs can be the reference's sampled list (P = 100 000) or all N^2 ordered pairs; the torch stages run in chunks of
``chunk`` pairs so the [P,141] logits fit (the reference's dense branch chunks too, models/model.py:96).
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import ref_gpu, ref_model


class _Stages:
    def __init__(self):
        self.ev = {}

    def __call__(self, name):
        st = self

        class _C:
            def __enter__(self):
                self.a, self.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                self.a.record()

            def __exit__(self, *exc):
                self.b.record()
                st.ev.setdefault(name, []).append((self.a, self.b))
        return _C()

    def ms(self):
        torch.cuda.synchronize()
        return {k: sum(a.elapsed_time(b) for a, b in v) for k, v in self.ev.items()}


@torch.no_grad()
def estimate(pc, nrm, sd_pe, sd_ppf, idxs, cfg, sphere, seed=0, inject_tr=None, chunk=1 << 20):
    """pc, nrm: float32 CUDA [N,3]; sd_*: state_dicts on the same device; idxs: int32 CUDA [P,2].  inject_tr: optional float32
    CUDA [P,2] (mu, nu) used INSTEAD of the first-pass draws (the bench's trained-like vote load).  sphere: float32 CUDA
    [480,3].  -> dict(stage_ms, flat, n_survivors, best, wall_ms)."""
    dev = pc.device
    st = _Stages()
    t0 = time.perf_counter()
    torch.cuda.synchronize()
    n, p = pc.shape[0], idxs.shape[0]
    B, RB = cfg["tr_num_bins"], cfg["rot_num_bins"]
    g = torch.Generator(device=dev).manual_seed(int(seed))
    with st("point_encoder"):                                                      # nocs/inference.py:180-181
        dist = torch.cdist(pc[None], pc[None])[0]
        feat = ref_model.point_encode(pc, nrm, dist, sd_pe, cfg["knn"])
    with st("encode"):                                                             # :182-188
        tr = torch.empty((p, 2), dtype=torch.float32, device=dev)
        for c0 in range(0, p, chunk):
            logits = ref_model.ppf_encode_idx(pc, nrm, feat, idxs[c0:c0 + chunk], sd_ppf)
            pr = torch.softmax(logits[:, :2 * B].reshape(-1, 2, B), -1)
            b = torch.cat([torch.multinomial(pr[:, 0], 1, generator=g), torch.multinomial(pr[:, 1], 1, generator=g)], -1)
            tr[c0:c0 + chunk] = ref_model.decode_tr(b[:, 0], b[:, 1], B, cfg["vote_range"])
            del logits, pr
    if inject_tr is not None:
        tr = inject_tr
    corner = pc.min(0)[0]
    dims = tuple(int(v) for v in (torch.div(pc.max(0)[0] - corner, torch.full_like(corner, cfg["res"])).int() + 1).cpu())   # :194-195
    probs = torch.ones(n, device=dev)
    grid = torch.zeros(dims, dtype=torch.float32, device=dev)
    with st("vote"):                                                               # :197-205, the reference's over-launch (:192)
        ref_gpu.ppf_voting(pc, tr, probs, idxs, grid, corner, cfg["res"], 72, True, over_launch=(p <= n * n))
    with st("argmax"):                                                             # :207-211 grid .get() + np.argmax
        flat = int(np.argmax(grid.cpu().numpy(), axis=None))
    cell = np.array(np.unravel_index(flat, dims))
    centre = torch.from_numpy((corner.double().cpu().numpy() + cell * cfg["res"]).astype(np.float32)).to(dev)
    with st("backvote"):                                                           # :216-231
        oc = torch.zeros((p, 3), dtype=torch.float32, device=dev)
        ref_gpu.backvote(pc, tr, oc, idxs, corner, cfg["res"], 72, dims, centre, 3 * cfg["res"])
        mask = (oc != 0).any(-1)
        kept = idxs[mask]
    m = kept.shape[0]
    out = {"flat": flat, "n_survivors": int(m), "dims": dims}
    if m:
        with st("encode2"):                                                        # :236-256
            rot = torch.empty(m, dtype=torch.float32, device=dev)
            aux = torch.empty(m, dtype=torch.float32, device=dev)
            ls = torch.zeros(3, dtype=torch.float64, device=dev)
            for c0 in range(0, m, chunk):
                l2 = ref_model.ppf_encode_idx(pc, nrm, feat, kept[c0:c0 + chunk], sd_ppf)
                up = torch.multinomial(torch.softmax(l2[:, 2 * B:2 * B + RB], -1), 1, generator=g)[:, 0]
                rot[c0:c0 + chunk] = ref_model.decode_rot(up, RB)
                aux[c0:c0 + chunk] = l2[:, -5]
                ls += l2[:, -3:].double().sum(0)
                del l2
        with st("rot_vote"):                                                       # :258-284: ALL survivors, then 10 000 of them
            cand = torch.zeros((m, 72, 3), dtype=torch.float32, device=dev)
            ref_gpu.rot_voting(pc, rot, cand, kept, 72)
            sel = torch.randperm(m, generator=g, device=dev)[:10000]
            cos = cand[sel].reshape(-1, 3) @ sphere.T
            counts = (cos > float(np.cos(cfg.get("angle_prec", 1.5) / 180 * np.pi))).sum(0)
            best = int(torch.argmax(counts).item())
        with st("aux_sign"):                                                       # :286-302 (numpy on the host upstream)
            ab = pc[kept[:, 0].long()] - pc[kept[:, 1].long()]
            abn = ab / (ab.pow(2).sum(-1).sqrt() + 1e-7)[:, None]
            pn = nrm[kept[:, 0].long()]
            pn = torch.where(((pn * abn).sum(-1) < 0)[:, None], -pn, pn)
            target = ((pn * sphere[best]).sum(-1) > 0).float()
            bce = torch.nn.functional.binary_cross_entropy_with_logits
            flip = bool(bce(aux, 1.0 - target).item() < bce(aux, target).item())
        out.update(best=best, flip=flip, log_scale=(ls / m).cpu().numpy())
    out["stage_ms"] = st.ms()
    out["wall_ms"] = (time.perf_counter() - t0) * 1e3
    return out
