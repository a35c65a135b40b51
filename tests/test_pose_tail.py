"""Host tail of the per-object path (cppf_b200/pipeline.py:_pose_from_record, nocs/inference.py:305-339): the pose built
from the small device record equals the oracle's restatement of those lines (oracle/ref_model.py:assemble_pose) --
centre from the argmax cell, sign-resolved up / right axes, Gram-Schmidt, column order by z_right, scale from the mean
log-scale.  Runs without a GPU (the estimator's constants live on the CPU here; no kernel is called)."""
import numpy as np
import pytest

from cppf_b200 import synth
from cppf_b200.pipeline import PoseConfig, PoseEstimator
from oracle import ref_model


@pytest.mark.parametrize("regress_right,z_right", [(False, False), (True, False), (True, True), (False, True)])
def test_pose_from_record_matches_the_restated_reference_tail(regress_right, z_right):
    rng = np.random.default_rng(5 + 2 * regress_right + z_right)
    cfg = PoseConfig.from_dict(dict(synth.BOTTLE, regress_right=regress_right, z_right=z_right))
    est = PoseEstimator(None, None, cfg, "cpu")
    n_dirs = 2 if regress_right else 1
    dims = (18, 55, 18)
    for _ in range(50):
        flat = int(rng.integers(0, dims[0] * dims[1] * dims[2]))
        bests = [int(rng.integers(0, est.sphere_np.shape[0])) for _ in range(n_dirs)]
        count = float(rng.integers(1, 5000))
        log_sum = rng.normal(0, 0.3, 3) * count
        s_up, s_right = rng.normal(0, 5.0, 2)
        corner = rng.normal(0, 0.1, 3).astype(np.float32).astype(np.float64)
        rec = np.concatenate([[flat], bests, log_sum, [count, s_up, s_right], corner])
        out = est._pose_from_record(rec, n_dirs, dims)
        cell = np.array(np.unravel_index(flat, dims))
        T = corner + cell * cfg.res                                             # nocs/inference.py:209
        up = est.sphere_np[bests[0]] * (-1.0 if s_up < 0 else 1.0)              # :299-302
        right = est.sphere_np[bests[1]] * (-1.0 if s_right < 0 else 1.0) if regress_right else None
        RT, scales = ref_model.assemble_pose(up, right, T, (log_sum / count).astype(np.float32), cfg.scale_mean,
                                             z_right=z_right, regress_right=regress_right, scale_mul=cfg.scale_mul)
        np.testing.assert_allclose(out["RT"], RT, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(out["scales"], scales, rtol=1e-6)
        assert out["argmax_flat"] == flat and out["n_survivors"] == int(count)
        R = out["RT"][:3, :3] / np.linalg.norm(out["pred_scale"])
        np.testing.assert_allclose(R.T @ R, np.eye(3), atol=1e-5)               # a rotation (up to reflection), as in the reference


def test_zero_survivors_raise_and_dims_travel_with_the_pending_pose():
    """a record with no survivors must not come back as a normal-looking pose, and the dims a flat argmax is
    unravelled with belong to the pending pose, not to the estimator (two staged objects of one category in flight)."""
    import torch
    from cppf_b200.pipeline import NoSurvivorsError, PendingPose
    cfg = PoseConfig.from_dict(dict(synth.BOTTLE))
    est = PoseEstimator(None, None, cfg, "cpu")
    corner = np.zeros(3)
    rec = np.concatenate([[7.0], [3.0], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0], corner])
    with pytest.raises(NoSurvivorsError):
        est._pose_from_record(rec, 1, (4, 5, 6))

    class _Done:
        def synchronize(self):
            pass
    rec[5] = 10.0                                   # 10 survivors
    flat = 4 * 5 * 6 - 1
    rec[0] = flat
    a = PendingPose(est, torch.from_numpy(rec.copy()), _Done(), 1, staged=True, dims=(4, 5, 6))
    b = PendingPose(est, torch.from_numpy(rec.copy()), _Done(), 1, staged=True, dims=(6, 5, 4))
    ra, rb = b.result(), a.result()                 # read in the "wrong" order on purpose
    np.testing.assert_allclose(rb["T_host"], np.array([3, 4, 5]) * cfg.res)
    np.testing.assert_allclose(ra["T_host"], np.array([5, 4, 3]) * cfg.res)
