"""TEST INFRASTRUCTURE (oracle) -- mint golden fixtures by RUNNING THE REFERENCE.

The reference has no tests and no golden vectors, so the only way to
pin parity is to execute its code on seeded inputs and commit the input/output
pairs.  This script (run in the build container, where /root/reference exists):

* imports ``models/model.py`` + ``models/sprin.py`` unmodified (torch CPU) ->
  ``tests/golden/encoder_bottle.npz``  (PointEncoder / PPFEncoder, indexed + dense);
* runs the reference's own voting kernel strings, compiled for the CPU by
  ``oracle/build_ref.py`` -> ``tests/golden/voting_bottle.npz``;
* ``ast``-extracts ``fibonacci_sphere`` (utils/util.py:102-118) and ``generate_target``
  (utils/dataset.py:20-60) -- their modules cannot be imported (open3d, pyrender) --
  and executes them -> ``tests/golden/host_glue.npz``;
* records ``torch.multinomial`` draws next to the Exp(1) noise that reproduces them;
* ``ast``-extracts ``backproject`` (utils/util.py:598-631) and runs it on a window of the reference's demo depth
  frame -> ``tests/golden/preprocess_demo.npz``.

Nothing here is needed at test time on the GPU box; the .npz files travel.
"""
from __future__ import annotations

import ast
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("CPPF_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from cppf_b200 import synth  # noqa: E402  (pure numpy helpers)
from oracle import clib  # noqa: E402


def _extract_function(path, name):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            return ast.get_source_segment(open(path).read(), node)
    raise KeyError(name)


def encoder_fixture():
    sys.path.insert(0, REF)
    from models.model import PPFEncoder, PointEncoder  # the reference itself

    torch.manual_seed(0)
    point_encoder = PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).eval()
    ppf_encoder = PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=2 * 32 + 2 * 36 + 2 + 3).eval()
    n, p, nd = 256, 512, 24
    pc, nrm = synth.synth_bottle(n, 0)
    idxs = synth.sample_pairs(n, p, 0)
    with torch.no_grad():
        pcs, nrms = torch.from_numpy(pc)[None], torch.from_numpy(nrm)[None]
        dist = torch.cdist(pcs, pcs)                                   # nocs/inference.py:180
        nbrs = torch.topk(dist, 60, largest=False, sorted=False)[1]     # models/model.py:47
        feat = point_encoder(pcs, nrms, dist)                          # :181
        feat_nbrs = point_encoder.forward_nbrs(pcs, nrms, nbrs)
        logits = ppf_encoder(pcs, nrms, feat, idxs=idxs)               # :182
        # dense branch on the first nd points with an exact distance matrix (SURVEY 8a notes)
        sub = slice(0, nd)
        dist_exact = (pcs[:, sub, None] - pcs[:, None, sub]).norm(dim=-1)
        dense = ppf_encoder(pcs[:, sub], nrms[:, sub], feat[:, sub], dist=dist_exact)
    out = dict(pc=pc, nrm=nrm, idxs=idxs, nbrs=nbrs[0].numpy().astype(np.int64), feat=feat[0].numpy(),
               feat_nbrs=feat_nbrs[0].numpy(), logits=logits[0].numpy(), dense_n=np.int64(nd),
               dense_dist=dist_exact[0].numpy(), dense_logits=dense[0].numpy())
    for k, v in point_encoder.state_dict().items():
        out["pe/" + k] = v.numpy()
    for k, v in ppf_encoder.state_dict().items():
        out["ppf/" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "encoder_bottle.npz"), **out)
    print("encoder_bottle.npz", {k: v.shape for k, v in out.items() if "/" not in k})


def voting_fixture():
    assert clib.have_ref_cpu(), "run oracle/build_ref.py first"
    cfg = synth.BOTTLE
    n, p = 512, 4096
    pc, nrm = synth.synth_bottle(n, 1)
    idxs = synth.sample_pairs(n, p, 1).astype(np.int32)
    idxs[:8, 1] = idxs[:8, 0]                                           # degenerate pairs (voting.py:21)
    tr = synth.trained_like_tr(pc, idxs, cfg["tr_num_bins"], cfg["vote_range"])
    corner, dims = synth.vote_grid_geometry(pc, cfg["res"])
    probs = np.ones(n, np.float32)                                      # nocs/inference.py:201
    out = dict(pc=pc, nrm=nrm, idxs=idxs, tr=tr, corner=corner, dims=np.array(dims, np.int64),
               res=np.float32(cfg["res"]))
    for adaptive in (True, False):
        g = clib.ppf_voting(pc, tr, probs, idxs, dims, corner, cfg["res"], 72, adaptive, impl="ref_cpu")
        out[f"grid_adaptive{int(adaptive)}"] = g
        out[f"argmax_adaptive{int(adaptive)}"] = np.int64(np.argmax(g))
    # random-bin targets (mostly out-of-grid votes) + non-uniform probs exercise the other branches
    rng = np.random.default_rng(7)
    tr_rand = np.stack([rng.integers(0, 32, p) / 31 * 0.5 - 0.25, rng.integers(0, 32, p) / 31 * 0.25], -1).astype(np.float32)
    probs_rand = rng.uniform(0.1, 1.0, n).astype(np.float32)
    out["tr_rand"], out["probs_rand"] = tr_rand, probs_rand
    out["grid_rand"] = clib.ppf_voting(pc, tr_rand, probs_rand, idxs, dims, corner, cfg["res"], 72, True, impl="ref_cpu")
    flat = int(np.argmax(out["grid_adaptive1"]))
    centre = (corner.astype(np.float64) + np.array(np.unravel_index(flat, dims)) * cfg["res"]).astype(np.float32)
    out["centre"] = centre
    out["backvote"] = clib.backvote(pc, tr, idxs, dims, corner, cfg["res"], centre, 3 * cfg["res"], 72, impl="ref_cpu")
    rot = synth.trained_like_rot(pc, idxs[:96], cfg["rot_num_bins"], cfg["up_sym"])
    out["rot"] = rot
    out["rot_candidates"] = clib.rot_voting(pc, rot, idxs[:96], 72, impl="ref_cpu")
    out["findpeak_w1"] = clib.findpeak(out["grid_adaptive1"], 1, literal=True, impl="ref_cpu")
    out["findpeak_w2"] = clib.findpeak(out["grid_adaptive1"], 2, literal=True, impl="ref_cpu")
    np.savez_compressed(os.path.join(OUT, "voting_bottle.npz"), **out)
    print("voting_bottle.npz dims", dims, "argmax", flat, "survivors", int(np.any(out["backvote"] != 0, -1).sum()))


def host_glue_fixture():
    ns = {"math": math, "np": np}
    exec(_extract_function(os.path.join(REF, "utils", "util.py"), "fibonacci_sphere"), ns)
    exec(_extract_function(os.path.join(REF, "utils", "dataset.py"), "generate_target"), ns)
    sphere = np.array(ns["fibonacci_sphere"](int(4 * np.pi / (1.5 / 180 * np.pi))))    # nocs/inference.py:100-102
    pc, nrm = synth.synth_bottle(128, 2)
    np.random.seed(0)
    t_tr, t_rot, t_aux, t_idx = ns["generate_target"](pc.astype(np.float64), nrm.astype(np.float64),
                                                     up_sym=True, subsample=1000)
    # torch.multinomial == argmax(p/q), q ~ Exp(1) from the same generator state
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(2000, 32, generator=g) * 3
    probs = torch.softmax(logits, -1)
    g1 = torch.Generator().manual_seed(11)
    draws = torch.multinomial(probs, 1, generator=g1)[:, 0]
    g2 = torch.Generator().manual_seed(11)
    q = torch.empty_like(probs).exponential_(1, generator=g2)
    np.savez_compressed(os.path.join(OUT, "host_glue.npz"), sphere=sphere, pc=pc, nrm=nrm, target_tr=t_tr,
                        target_rot=t_rot, target_aux=t_aux, target_idx=t_idx, mn_logits=logits.numpy(),
                        mn_draws=draws.numpy(), mn_q=q.numpy())
    print("host_glue.npz sphere", sphere.shape, "multinomial==race:",
          bool((torch.argmax(probs / q, -1) == draws).all()))


def preprocess_fixture():
    """utils/util.py:598-631 `backproject` (ast-extracted: the module imports open3d) executed on a window of the
    reference's demo frame data/demo/0000_depth.png with the NOCS intrinsics of nocs/inference.py:98."""
    import cv2
    ns = {"np": np}
    exec(_extract_function(os.path.join(REF, "utils", "util.py"), "backproject"), ns)
    depth_full = cv2.imread(os.path.join(REF, "data", "demo", "0000_depth.png"), -1)
    r0, c0, h, w = 200, 260, 96, 128
    depth = np.ascontiguousarray(depth_full[r0:r0 + h, c0:c0 + w])
    yy, xx = np.mgrid[0:h, 0:w]
    mask = ((yy - 48) ** 2 / 40.0 ** 2 + (xx - 64) ** 2 / 56.0 ** 2) <= 1.0           # an elliptical "instance"
    intr = np.array([[591.0125, 0, 322.525], [0, 590.16775, 244.11084], [0, 0, 1]])    # nocs/inference.py:98
    intr_win = intr.copy()
    intr_win[0, 2] -= c0
    intr_win[1, 2] -= r0
    pts, idxs = ns["backproject"](depth, intr_win, mask)
    np.savez_compressed(os.path.join(OUT, "preprocess_demo.npz"), depth=depth, mask=mask, intrinsics=intr_win, pts=pts,
                        rows=idxs[0].astype(np.int64), cols=idxs[1].astype(np.int64))
    print("preprocess_demo.npz", depth.shape, depth.dtype, "valid", len(pts))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)            # deterministic reduction order for the fixtures
    only = sys.argv[1:]
    for name, fn in (("encoder", encoder_fixture), ("voting", voting_fixture), ("host_glue", host_glue_fixture),
                     ("preprocess", preprocess_fixture)):
        if not only or name in only:
            fn()
