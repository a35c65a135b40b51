"""CPU: the pre-processing oracle (oracle/ref_preprocess.py) against the fixture minted by running the reference's own
`backproject` on its demo depth frame (oracle/make_golden.py preprocess), plus the stated semantics of the
MinkowskiEngine / open3d stand-ins."""
import os

import numpy as np

from oracle import ref_preprocess as rp

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess_demo.npz"))


def test_backproject_restatement_equals_reference_output():
    pts, idxs = rp.backproject(GOLD["depth"], GOLD["intrinsics"], GOLD["mask"])
    np.testing.assert_array_equal(idxs[0], GOLD["rows"])
    np.testing.assert_array_equal(idxs[1], GOLD["cols"])
    np.testing.assert_array_equal(pts, GOLD["pts"])                     # same numpy expressions -> same bits
    pc, _ = rp.object_cloud(GOLD["depth"], GOLD["intrinsics"], GOLD["mask"])
    # nocs/inference.py:132,136-137: metres, axis flips undone
    np.testing.assert_array_equal(pc[:, 2], GOLD["pts"][:, 2] / 1000)
    np.testing.assert_array_equal(pc[:, 0], -(GOLD["pts"][:, 0] / 1000))


def test_sparse_quantize_keeps_first_point_of_every_voxel():
    rng = np.random.default_rng(0)
    pts = rng.uniform(-0.05, 0.05, (5000, 3))
    idx = rp.sparse_quantize(pts, 4e-3)
    vox = np.floor(pts / 4e-3).astype(np.int32)
    assert np.all(np.diff(idx) > 0)
    assert len(np.unique(vox[idx], axis=0)) == len(idx) == len(np.unique(vox, axis=0))
    seen = {}
    for i, v in enumerate(map(tuple, vox)):
        seen.setdefault(v, i)
    assert sorted(seen.values()) == idx.tolist()


def test_estimate_normals_of_a_plane_and_a_sphere():
    rng = np.random.default_rng(1)
    plane = np.c_[rng.uniform(-1, 1, (400, 2)), np.zeros(400)] @ np.linalg.qr(rng.normal(size=(3, 3)))[0].T
    n = rp.estimate_normals(plane.astype(np.float32), 20)
    axis = np.linalg.qr(rng.normal(size=(3, 3)))[0]      # not used: the plane normal is the third column of the rotation
    ref = np.cross(plane[1] - plane[0], plane[2] - plane[0])
    ref /= np.linalg.norm(ref)
    assert np.all(np.abs(n @ ref) > 1 - 1e-6)
    sph = rng.normal(size=(3000, 3))
    sph /= np.linalg.norm(sph, axis=1, keepdims=True)
    n = rp.estimate_normals(sph.astype(np.float32), 30)
    assert np.median(np.abs(np.sum(n * sph, -1))) > 0.99
