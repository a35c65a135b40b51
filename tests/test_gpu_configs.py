"""BASELINE.json configs 3-5 as parity cases at sizes that run in seconds (the full sizes are
tools/config_sweep.py): 64^3 vote grids with several categories batched, SUN RGB-D-like constants with
both orientation heads on dense pairs, and the dense stress shape."""
import numpy as np
import pytest
import torch

from cppf_b200 import model, synth
from cppf_b200.pipeline import PoseConfig, PoseEstimator, estimate_many

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _est(seed, cfg):
    torch.manual_seed(seed)
    pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(DEV).eval()
    ppf = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).to(DEV).eval()
    return PoseEstimator(pe, ppf, cfg, DEV)


def test_config3_categories_batched_on_64_cube_grids():
    """Six weight sets ("categories", nocs/inference.py:127-129) on clouds whose vote grid is exactly 64^3:
    too large for one SM's shared memory, so the vote runs on the global-reduction kernel; the batch API must
    return what the per-object calls return, from host and from device-resident clouds."""
    cfg = PoseConfig.from_dict(dict(synth.BOTTLE, n_pairs=40000, rot_subsample=2000))
    ests = [_est(s, cfg) for s in range(3)]
    clouds = [synth.synth_cylinder_grid64(1024, s) for s in range(3)]
    for p, _ in clouds:
        assert synth.vote_grid_geometry(p, cfg.res)[1] == (64, 64, 64)
    single = [e.estimate_fused(p, q, seed=s) for s, (e, (p, q)) in enumerate(zip(ests, clouds))]
    many_host = estimate_many([(e, p, q, s) for s, (e, (p, q)) in enumerate(zip(ests, clouds))])
    many_dev = estimate_many([(e, torch.from_numpy(p).to(DEV), torch.from_numpy(q).to(DEV), s)
                              for s, (e, (p, q)) in enumerate(zip(ests, clouds))])
    for a, b, c in zip(single, many_host, many_dev):
        assert a["argmax_flat"] == b["argmax_flat"] == c["argmax_flat"]
        assert a["best_bins"] == b["best_bins"] == c["best_bins"]
        assert a["n_survivors"] == b["n_survivors"] == c["n_survivors"] > 0
        np.testing.assert_array_equal(a["RT"], b["RT"])
    assert len({r["argmax_flat"] for r in single}) > 1          # different weights -> different votes


def test_config4_sunrgbd_constants_dense_pairs_both_heads():
    cfg = PoseConfig.from_dict(dict(synth.CHAIR, n_pairs=0, scale_mul=1.0, rot_subsample=5000))
    est = _est(1, cfg)
    pc, nrm = synth.synth_bottle(768, 3, scale=8.0)
    u = torch.rand(768 * 768, 4, generator=torch.Generator().manual_seed(5)).to(DEV)
    one = est.estimate_fused(pc, nrm, seed=2, uniforms=u)
    staged = est.estimate_fused(pc, nrm, seed=2, uniforms=u, staged=True)
    assert one["argmax_flat"] == staged["argmax_flat"] and one["best_bins"] == staged["best_bins"]
    assert len(one["best_bins"]) == 2 and one["n_survivors"] == staged["n_survivors"] > 0
    np.testing.assert_array_equal(one["RT"], staged["RT"])
    R = one["RT"][:3, :3] / np.linalg.norm(one["pred_scale"])
    np.testing.assert_allclose(R.T @ R, np.eye(3), atol=1e-5)       # Gram-Schmidt of nocs/inference.py:305-312
    assert abs(np.linalg.det(R) - 1) < 1e-5


@pytest.mark.parametrize("n", [1536, 2048])
def test_config5_dense_stress_shape(n):
    cfg = PoseConfig.from_dict(dict(synth.BOTTLE, n_pairs=0))
    est = _est(0, cfg)
    pc, nrm = synth.synth_bottle(n, 4)
    pcd = torch.from_numpy(pc).to(DEV)
    inj = synth.trained_like_bins_dense_torch(pcd, synth.BOTTLE)
    one = est.estimate_fused(pcd, torch.from_numpy(nrm).to(DEV), seed=1, inject_bins=inj)
    staged = est.estimate_fused(pc, nrm, seed=1, inject_bins=inj, staged=True)
    assert one["argmax_flat"] == staged["argmax_flat"] and one["n_survivors"] == staged["n_survivors"]
    # trained-like votes of a centred object: the winning cell is the one containing the origin
    corner, dims = synth.vote_grid_geometry(pc, cfg.res)
    cell = np.array(np.unravel_index(one["argmax_flat"], dims))
    assert np.all(np.abs(corner + cell * cfg.res) <= 1.5 * cfg.res)
