"""TEST INFRASTRUCTURE (oracle) -- mint a TRAINED checkpoint of the reference networks for the synthetic bottle.

The reference ships no checkpoints (``checkpoints/`` is git-ignored upstream) and a random-init pair
network votes at random: its vote grid has no peak, a handful of pairs survive the back-vote, and an end-to-end pose
comparison would hinge on near-ties.  This script (run in the build container, where /root/reference exists) imports the
REFERENCE modules ``models/model.py`` + ``models/sprin.py`` unmodified and trains them for a few minutes on the CPU with the
reference's own objective (``train.py:60-94``: KL divergence of the binned translation / rotation heads against the
two-bin soft targets of ``utils/util.py:121-146``, BCE on the aux head, MSE on the log-scale) on synthetic bottles
(``cppf_b200.synth.synth_bottle``, the benchmark's object), with ground truth from ``utils/dataset.py:20-60`` restated in
``oracle/ref_model.py`` (pinned to the reference function by tests/test_oracle_golden.py).

Output: ``tests/golden/trained_bottle.npz`` (``pe/*`` and ``ppf/*`` state-dict arrays under the reference's keys, ~90 KB).
Used by the end-to-end parity tests and by bench.py's ``variant_trained_network`` leg.  Deterministic for a fixed torch build
(``torch.manual_seed``); the committed file is the artefact, this script documents how it was made.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("CPPF_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from cppf_b200 import synth  # noqa: E402


def soft_bins(val, max_val, num_bins):
    """utils/util.py:121-146 (real2prob, non-circular): linear interpolation onto the two neighbouring bin centres."""
    interval = max_val / (num_bins - 1)
    low = torch.clamp(torch.floor(val / interval).long(), min=0, max=num_bins - 2)
    w_low = (1.0 - (val / interval - low)).clamp(0.0, 1.0)
    out = torch.zeros((*val.shape, num_bins), dtype=torch.float32)
    out.scatter_(-1, low[..., None], w_low[..., None])
    out.scatter_(-1, (low + 1)[..., None], (1.0 - w_low)[..., None])
    return out


def targets(pc, nrm, idxs, cfg):
    """utils/dataset.py:20-60 for a y-up object centred at the origin (up_sym as configured)."""
    a, b = pc[idxs[:, 0]].double(), pc[idxs[:, 1]].double()
    d = a - b
    du = d / (d.norm(dim=-1, keepdim=True) + 1e-7)
    mu = (a * du).sum(-1)
    nu = (a - mu[:, None] * du).norm(dim=-1)
    up = torch.arccos(du[:, 1].clamp(-1, 1))
    if cfg["up_sym"]:
        up = torch.minimum(up, torch.arccos((-du[:, 1]).clamp(-1, 1)))
    pn = nrm[idxs[:, 0]].double().clone()
    pn[(pn * du).sum(-1) < 0] *= -1
    aux = (pn[:, 1] > 0).float()
    return mu.float(), nu.float(), up.float(), aux


def main(steps=1800, n=1024, p=20000, out=os.path.join(ROOT, "tests", "golden", "trained_bottle.npz")):
    from models.model import PPFEncoder, PointEncoder  # the reference itself

    cfg = synth.BOTTLE
    B, RB = cfg["tr_num_bins"], cfg["rot_num_bins"]
    vr = cfg["vote_range"]
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    pe = PointEncoder(k=cfg["knn"], spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).train()
    ppf = PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=2 * B + 2 * RB + 2 + 3).train()
    opt = torch.optim.Adam([*pe.parameters(), *ppf.parameters()], lr=2e-3)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, steps, eta_min=1e-4)
    kld = torch.nn.KLDivLoss(reduction="batchmean")
    # extents of synth_bottle (radius 0.035 body, y in [-0.11, 0.11]) against scale_mean * 2 (nocs/inference.py:335)
    ext = torch.tensor([0.07, 0.22, 0.07])
    log_scale_t = torch.log(ext / (torch.tensor(cfg["scale_mean"]) * 2.0))
    t0 = time.time()
    for it in range(steps):
        pc_np, nrm_np = synth.synth_bottle(n, 100000 + it)
        pc, nrm = torch.from_numpy(pc_np), torch.from_numpy(nrm_np)
        idxs = torch.from_numpy(synth.sample_pairs(n, p, 100000 + it))
        mu, nu, up, aux = targets(pc, nrm, idxs, cfg)
        t_mu = soft_bins(mu + vr[0], 2 * vr[0], B)
        t_nu = soft_bins(nu, vr[1], B)
        t_up = soft_bins(up, float(np.pi), RB)
        opt.zero_grad()
        with torch.no_grad():
            dist = torch.cdist(pc[None], pc[None])
        feat = pe(pc[None], nrm[None], dist)
        preds = ppf(pc[None], nrm[None], feat, idxs=idxs)[0]
        loss_tr = kld(F.log_softmax(preds[:, :B], -1), t_mu) + kld(F.log_softmax(preds[:, B:2 * B], -1), t_nu)
        loss_up = kld(F.log_softmax(preds[:, 2 * B:2 * B + RB], -1), t_up)
        loss_aux = F.binary_cross_entropy_with_logits(preds[:, -5], aux)
        loss_scale = F.mse_loss(preds[:, -3:], log_scale_t[None].expand(p, 3))
        loss = loss_tr + loss_up + loss_aux + loss_scale
        loss.backward()
        opt.step()
        sched.step()
        if it % 100 == 0 or it == steps - 1:
            print(f"step {it:5d}  loss {loss.item():.4f}  tr {loss_tr.item():.4f}  up {loss_up.item():.4f}  "
                  f"aux {loss_aux.item():.4f}  scale {loss_scale.item():.5f}  [{time.time() - t0:.0f} s]", flush=True)
    arrays = {f"pe/{k}": v.detach().numpy().astype(np.float32) for k, v in pe.state_dict().items()}
    arrays.update({f"ppf/{k}": v.detach().numpy().astype(np.float32) for k, v in ppf.state_dict().items()})
    arrays["meta_steps"] = np.int64(steps)
    np.savez_compressed(out, **arrays)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main(steps=int(sys.argv[1]) if len(sys.argv) > 1 else 1800)
