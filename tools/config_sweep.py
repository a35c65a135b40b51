#!/usr/bin/env python
"""BASELINE.json configs 3-5 through the public API on one GPU (parity-test cases, not bench lines):
  3: six categories (six weight sets) batched, N=4096, 64^3 vote grids
  4: SUN RGB-D-like chair constants (res 3e-2, up + right heads), N=8192 dense pairs
  5: stress sweep N in {2048, 4096, 8192, 16384} dense pairs, bottle constants
Prints one JSON object; every timing is CUDA-event time over `reps` objects after one warm-up."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cppf_b200 import model, synth                                   # noqa: E402
from cppf_b200.pipeline import PoseConfig, PoseEstimator, estimate_many, release_workspaces  # noqa: E402

dev = torch.device("cuda")


def make_est(seed, cfg):
    torch.manual_seed(seed)
    pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(dev).eval()
    ppf = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).to(dev).eval()
    return PoseEstimator(pe, ppf, cfg, dev)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = {}
which = sys.argv[1:] or ["3", "4", "5"]
if "5" in which:
    cfg = PoseConfig.from_dict(dict(synth.BOTTLE, n_pairs=0))
    est = make_est(0, cfg)
    rows = []
    for n in (2048, 4096, 8192, 16384):
        pc, nrm = synth.synth_bottle(n, 0)
        pcd, nd = torch.from_numpy(pc).to(dev), torch.from_numpy(nrm).to(dev)
        inj = synth.trained_like_bins_dense_torch(pcd, synth.BOTTLE, chunk_rows=max(16, 2 ** 20 // n))
        reps = 8 if n <= 8192 else 3
        pend = []
        ms = timed(lambda: est.enqueue_fused(pcd, nd, seed=0, inject_bins=inj).result(), reps)
        rows.append({"n_points": n, "pairs": n * n, "ms_per_object": ms, "pairs_per_s": n * n / ms * 1e3})
        del inj
        release_workspaces()
        torch.cuda.empty_cache()
    out["config5_stress_dense_bottle"] = rows
if "4" in which:
    cfg = PoseConfig.from_dict(dict(synth.CHAIR, n_pairs=0, scale_mul=1.0))
    est = make_est(1, cfg)
    n = 8192
    pc, nrm = synth.synth_bottle(n, 3, scale=8.0)
    pcd, nd = torch.from_numpy(pc).to(dev), torch.from_numpy(nrm).to(dev)
    res = {}
    ms = timed(lambda: res.update(est.enqueue_fused(pcd, nd, seed=0).result()), 5)
    out["config4_sunrgbd_like_dense_8192"] = {"ms_per_object": ms, "pairs_per_s": n * n / ms * 1e3, "grid_dims": list(res["grid_dims"]),
                                              "n_survivors": res["n_survivors"]}
    release_workspaces()
    torch.cuda.empty_cache()
if "3" in which:
    import ctypes as C
    from cppf_b200 import _lib
    L = _lib.lib()
    cfg = PoseConfig.from_dict(dict(synth.BOTTLE, n_pairs=0))
    ests = [make_est(s, cfg) for s in range(6)]
    n = 4096
    clouds = [synth.synth_cylinder_grid64(n, s) for s in range(6)]
    dclouds = [(torch.from_numpy(p).to(dev), torch.from_numpy(q).to(dev)) for p, q in clouds]
    names = [L.cppf_timing_stage_name(i).decode() for i in range(L.cppf_timing_stages())]
    variants = (("network_votes", False), ("trained_like_votes", True))
    if os.environ.get("CPPF_SWEEP_VOTES"):
        variants = tuple(v for v in variants if v[0].startswith(os.environ["CPPF_SWEEP_VOTES"]))
    for tag, inject in variants:
        injs = [synth.trained_like_bins_dense_torch(p, synth.BOTTLE) if inject else None for p, _ in dclouds]
        timing = L.cppf_timing_create()

        def batch():
            pend = [e.enqueue_fused(p, q, seed=s, inject_bins=injs[s], max_cells=1, routed_max_cells=64 ** 3)
                    for s, (e, (p, q)) in enumerate(zip(ests, dclouds))]
            last.update(pend[0].result())
            return [x.result() for x in pend[1:]]
        last = {}
        ms = timed(batch, 2)
        for e in ests:
            e.timing = timing
        batch()
        acc = (C.c_float * len(names))()
        calls = L.cppf_timing_collect(timing, acc)
        for e in ests:
            e.timing = None
        L.cppf_timing_destroy(timing)
        out["config3_six_categories_64cube_dense_4096_" + tag] = {
            "ms_per_batch": ms, "ms_per_object": ms / 6, "pairs_per_s": 6 * n * n / ms * 1e3, "grid_dims": list(last["grid_dims"]),
            "stage_ms_per_object": {nm: acc[i] / max(calls, 1) for i, nm in enumerate(names)}}
        del injs
print(json.dumps(out))
