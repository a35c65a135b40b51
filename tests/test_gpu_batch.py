"""GPU: cppf_pose_batch (the object loop of nocs/inference.py:120-129 as one library call: objects fanned out over worker
streams, launches issued by several host threads, pairs drawn on the device) produces, object for object, the SAME record
as one cppf_pose_fused call per object on one stream."""
import numpy as np
import pytest
import torch

from cppf_b200 import model, synth
from cppf_b200.pipeline import PoseConfig, PoseEstimator, enqueue_batch, estimate_many

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _est(seed, **over):
    torch.manual_seed(seed)
    pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(DEV).eval()
    ppf = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).to(DEV).eval()
    return PoseEstimator(pe, ppf, PoseConfig.from_dict(dict(synth.BOTTLE, **over)), DEV)


@pytest.mark.parametrize("n_streams,n_threads", [(1, 1), (3, 2), (4, 4)])
def test_batch_records_equal_per_object_calls(n_streams, n_threads):
    ests = [_est(0, n_pairs=20000, rot_subsample=1500), _est(1, n_pairs=30000, rot_subsample=0, regress_right=True),
            _est(2, n_pairs=0, rot_subsample=1000)]                        # the third category runs dense N^2 pairs
    sizes = [700, 300, 1500, 412, 256, 999, 640, 333, 1200, 128, 777, 500, 350]
    items = []
    for i, n in enumerate(sizes):
        est = ests[i % 3]
        if est.cfg.n_pairs == 0:
            n = min(n, 400)
        pc, nrm = synth.synth_bottle(n, 50 + i)
        pc = pc + np.float32([0.01 * i, -0.02 * i, 0.5])                     # objects sit at different places
        items.append((est, pc.astype(np.float32), nrm, 1000 + i))
    single = [est.enqueue_fused(pc, nrm, seed=seed, device_pairs=True).result() for est, pc, nrm, seed in items]
    pend = enqueue_batch(items, n_streams=n_streams, n_threads=n_threads)
    batch = pend.results()
    for a, b in zip(single, batch):
        np.testing.assert_array_equal(a["record"], b["record"])
        np.testing.assert_array_equal(a["RT"], b["RT"])
        assert a["n_survivors"] == b["n_survivors"] and a["argmax_flat"] == b["argmax_flat"] and a["grid_dims"] == b["grid_dims"]
    # the column-wise host tail equals the per-object one
    np.testing.assert_allclose(pend.records17(), np.stack([a["record"] for a in single]), rtol=1e-6, atol=1e-7)
    # CUDA-resident clouds take the same route
    dev_items = [(e, torch.from_numpy(p).to(DEV), torch.from_numpy(q).to(DEV), s) for e, p, q, s in items[:4]]
    for a, b in zip(single[:4], estimate_many(dev_items, n_streams=2, n_threads=2)):
        np.testing.assert_array_equal(a["record"], b["record"])


def test_device_drawn_pairs_are_uniform_and_seeded():
    """nocs/inference.py:177 draws P uniform pairs per object; the in-library draw is a function of (seed, pair index)."""
    est = _est(0, n_pairs=50000, rot_subsample=1000)
    pc, nrm = synth.synth_bottle(800, 3)
    a = est.enqueue_fused(pc, nrm, seed=7, device_pairs=True).result()
    b = est.enqueue_fused(pc, nrm, seed=7, device_pairs=True).result()
    c = est.enqueue_fused(pc, nrm, seed=8, device_pairs=True).result()
    np.testing.assert_array_equal(a["record"], b["record"])
    assert not np.array_equal(a["record"], c["record"])
