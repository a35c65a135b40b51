"""GPU: device pre-processing (cppf_backproject / cppf_voxel_first / cppf_normals_pca) against the
oracle and the reference-minted fixture, and the per-image driver on the reference's demo depth window."""
import os

import numpy as np
import pytest
import torch

from cppf_b200 import inference, model, preprocess, synth
from cppf_b200.pipeline import PoseConfig, PoseEstimator
from oracle import ref_preprocess as rp

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess_demo.npz"))


def _depth_t(a):
    return torch.from_numpy(a.astype(np.int32)).to(DEV).to(torch.uint16)


def test_backproject_matches_reference_fixture():
    depth, mask = GOLD["depth"], GOLD["mask"]
    pts, pix = preprocess.backproject(_depth_t(depth), torch.from_numpy(mask).to(DEV), GOLD["intrinsics"])
    np.testing.assert_array_equal(pix.cpu().numpy(), GOLD["rows"] * depth.shape[1] + GOLD["cols"])
    ref, _ = rp.object_cloud(depth, GOLD["intrinsics"], mask)
    np.testing.assert_allclose(pts.cpu().numpy(), ref, rtol=1e-13, atol=1e-15)
    # float32 depth, empty mask, all-zero depth
    pts_f, _ = preprocess.backproject(torch.from_numpy(depth.astype(np.float32)).to(DEV), torch.from_numpy(mask).to(DEV),
                                      GOLD["intrinsics"])
    np.testing.assert_allclose(pts_f.cpu().numpy(), ref, rtol=1e-13, atol=1e-15)
    assert preprocess.backproject(_depth_t(depth), torch.zeros_like(torch.from_numpy(mask)).to(DEV))[0].shape[0] == 0
    assert preprocess.backproject(_depth_t(np.zeros_like(depth)), torch.from_numpy(mask).to(DEV))[0].shape[0] == 0


@pytest.mark.parametrize("m,voxel", [(4716, 4e-3), (200000, 2e-3), (1, 1e-2)])
def test_voxel_first_matches_oracle(m, voxel):
    if m == 4716:
        pts = rp.object_cloud(GOLD["depth"], GOLD["intrinsics"], GOLD["mask"])[0]
    else:
        pts = np.random.default_rng(m).uniform(-0.2, 0.2, (m, 3))
        pts[m // 2:] = pts[:m - m // 2]                     # exact duplicates: ties go to the lower index
    pc, idx = preprocess.sparse_quantize(torch.from_numpy(pts).to(DEV), voxel)
    ref = rp.sparse_quantize(pts, voxel)
    np.testing.assert_array_equal(idx.cpu().numpy(), ref)
    np.testing.assert_array_equal(pc.cpu().numpy(), pts[ref].astype(np.float32))     # nocs/inference.py:141


def test_normals_match_oracle_up_to_sign():
    pc, nrm = synth.synth_bottle(1500, 3)
    pc = (pc + np.random.default_rng(0).normal(0, 2e-4, pc.shape)).astype(np.float32)
    got = preprocess.estimate_normals(torch.from_numpy(pc).to(DEV), 30).cpu().numpy()
    ref = rp.estimate_normals(pc, 30)
    np.testing.assert_allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)
    agree = np.abs(np.sum(got * ref, -1))
    assert np.mean(agree > 1 - 1e-4) > 0.995          # same eigenvector up to sign (isolated near-degenerate neighbourhoods aside)
    assert np.median(np.abs(np.sum(got * nrm, -1))) > 0.95          # and it is the surface normal
    toward = preprocess.estimate_normals(torch.from_numpy(pc + np.float32([0, 0, 1])).to(DEV), 30, orient=True).cpu().numpy()
    assert np.all(np.sum(toward * (pc + np.float32([0, 0, 1])), -1) <= 1e-6)
    # degenerate input: all points equal -> open3d's (0, 0, 1)
    same = torch.zeros(40, 3, device=DEV)
    np.testing.assert_array_equal(preprocess.estimate_normals(same, 10).cpu().numpy(), np.tile([0, 0, 1], (40, 1)).astype(np.float32))


def test_driver_on_the_reference_demo_window():
    torch.manual_seed(0)
    pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(DEV).eval()
    ppf = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).to(DEV).eval()
    cfg = PoseConfig.from_dict(dict(synth.BOTTLE, res=1e-2, n_pairs=20000))
    est = PoseEstimator(pe, ppf, cfg, DEV)
    masks = np.stack([GOLD["mask"], np.zeros_like(GOLD["mask"]), GOLD["mask"]], -1)
    out = inference.estimate_image(GOLD["depth"], masks, [1, 1, 1], {"bottle": est}, intrinsics=GOLD["intrinsics"], seed=3)
    assert out["pred_RTs"].shape == (3, 4, 4) and out["pred_scales"].shape == (3, 3)
    assert out["n_points"][1] == 0 and np.array_equal(out["pred_RTs"][1], np.eye(4, dtype=np.float32))      # empty mask: identity
    assert out["n_points"][0] >= cfg.knn and np.isfinite(out["pred_RTs"]).all()
    R = out["pred_RTs"][0][:3, :3]
    s = np.cbrt(abs(np.linalg.det(R)))
    np.testing.assert_allclose((R / s).T @ (R / s), np.eye(3), atol=1e-4)
    # the translation is a grid cell inside the object's bounding box (nocs/inference.py:209)
    pc, _ = rp.object_cloud(GOLD["depth"], GOLD["intrinsics"], GOLD["mask"])
    T = out["pred_RTs"][0][:3, 3]
    assert np.all(T >= pc.min(0) - 2 * cfg.res) and np.all(T <= pc.max(0) + 2 * cfg.res)
