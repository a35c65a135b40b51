"""The trained-like vote load bench.py runs under (cppf_b200/synth.py:trained_like_*) is the reference's own ground-truth
target (utils/dataset.py:27-45, pinned by the `generate_target` fixture minted from the reference function) snapped to the
bin centres of nocs/inference.py:187-188,252 -- checked here against that fixture.  CPU only."""
import os

import numpy as np
import torch

from cppf_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "host_glue.npz")


def test_trained_like_tr_is_the_reference_target_snapped_to_bin_centres():
    g = np.load(GOLD)
    pc, idx, target = g["pc"], g["target_idx"], g["target_tr"]              # target_tr: output of utils/dataset.py:27-36
    B, vr = 32, (0.25, 0.25)
    q = synth.trained_like_tr(pc, idx, B, vr)
    step_mu, step_nu = 2 * vr[0] / (B - 1), vr[1] / (B - 1)
    inside = (np.abs(target[:, 0]) <= vr[0]) & (target[:, 1] <= vr[1])
    assert inside.mean() > 0.9
    # nearest bin centre: within half a bin of the reference target, and exactly a bin centre
    assert np.all(np.abs(q[inside, 0] - target[inside, 0]) <= 0.5 * step_mu + 1e-6)
    assert np.all(np.abs(q[inside, 1] - target[inside, 1]) <= 0.5 * step_nu + 1e-6)
    b_mu = np.rint((q[:, 0] + vr[0]) / step_mu)
    b_nu = np.rint(q[:, 1] / step_nu)
    centres_mu = (b_mu.astype(np.float32) / np.float32(B - 1) * np.float32(2 * vr[0]) - np.float32(vr[0]))
    centres_nu = b_nu.astype(np.float32) / np.float32(B - 1) * np.float32(vr[1])
    np.testing.assert_array_equal(q[:, 0], centres_mu.astype(np.float32))
    np.testing.assert_array_equal(q[:, 1], centres_nu.astype(np.float32))


def test_trained_like_rot_is_the_reference_target_snapped_to_bin_centres():
    g = np.load(GOLD)
    pc, idx, target = g["pc"], g["target_idx"], g["target_rot"][:, 0]       # utils/dataset.py:38-45, up_sym = True
    R = 36
    q = synth.trained_like_rot(pc, idx, R, True)
    assert np.all(np.abs(q - target) <= 0.5 * np.pi / (R - 1) + 1e-5)
    b = np.rint(q / np.pi * (R - 1))
    np.testing.assert_allclose(q, b.astype(np.float32) / np.float32(R - 1) * np.float32(np.pi), rtol=0, atol=0)


def test_dense_device_bins_equal_the_indexed_host_targets():
    pc, _ = synth.synth_bottle(96, 3)
    n = pc.shape[0]
    bins = synth.trained_like_bins_dense_torch(torch.from_numpy(pc), synth.BOTTLE, chunk_rows=40).numpy()
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    idx = np.stack([ii, jj], -1).reshape(-1, 2)
    tr = synth.trained_like_tr(pc, idx, 32, synth.BOTTLE["vote_range"])
    rot = synth.trained_like_rot(pc, idx, 36, True)
    b_mu = np.rint((tr[:, 0] + 0.25) / 0.5 * 31).astype(np.uint8)
    b_nu = np.rint(tr[:, 1] / 0.25 * 31).astype(np.uint8)
    b_up = np.rint(rot / np.pi * 35).astype(np.uint8)
    off = idx[:, 0] != idx[:, 1]                                             # i = j pairs are dropped by every vote kernel
    # float32 cloud promoted to float64 on both sides: the same arithmetic, so the bins agree except on exact .5 ties
    assert (bins[off, 0] != b_mu[off]).mean() < 1e-3
    assert (bins[off, 1] != b_nu[off]).mean() < 1e-3
    assert (bins[off, 2] != b_up[off]).mean() < 1e-3
