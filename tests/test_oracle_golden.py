"""CPU: the oracle (oracle/) against fixtures minted from the reference itself
(tests/golden/, oracle/make_golden.py) and, where present here, against the
reference's own kernel strings compiled for the CPU (oracle/_ref)."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden, split_state
from oracle import clib, ref_model
from parity_util import assert_masks_equal_up_to_rounding


@pytest.fixture(scope="module")
def enc():
    return load_golden("encoder_bottle.npz")


@pytest.fixture(scope="module")
def vot():
    return load_golden("voting_bottle.npz")


@pytest.fixture(scope="module")
def glue():
    return load_golden("host_glue.npz")


def test_pair_mlp_indexed_matches_reference(enc):
    sd = split_state(enc, "ppf/")
    out = ref_model.ppf_encode_idx(torch.from_numpy(enc["pc"]), torch.from_numpy(enc["nrm"]),
                                   torch.from_numpy(enc["feat"]), enc["idxs"], sd)
    np.testing.assert_allclose(out.numpy(), enc["logits"], rtol=1e-5, atol=2e-6)


def test_pair_mlp_dense_matches_reference(enc):
    sd = split_state(enc, "ppf/")
    n = int(enc["dense_n"])
    out = ref_model.ppf_encode_dense(torch.from_numpy(enc["pc"][:n]), torch.from_numpy(enc["nrm"][:n]),
                                     torch.from_numpy(enc["feat"][:n]), torch.from_numpy(enc["dense_dist"]), sd)
    np.testing.assert_allclose(out.numpy(), enc["dense_logits"], rtol=1e-5, atol=2e-6)


def test_point_encoder_matches_reference(enc):
    sd = split_state(enc, "pe/")
    pc, nrm = torch.from_numpy(enc["pc"]), torch.from_numpy(enc["nrm"])
    feat = ref_model.point_encode_nbrs(pc, nrm, torch.from_numpy(enc["nbrs"]), sd)
    np.testing.assert_allclose(feat.numpy(), enc["feat_nbrs"], rtol=1e-4, atol=1e-5)
    dist = torch.cdist(pc[None], pc[None])[0]
    feat2 = ref_model.point_encode(pc, nrm, dist, sd, 60)
    np.testing.assert_allclose(feat2.numpy(), enc["feat"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("adaptive", [True, False])
def test_vote_oracle_matches_reference_kernel(vot, adaptive):
    probs = np.ones(vot["pc"].shape[0], np.float32)
    g = clib.ppf_voting(vot["pc"], vot["tr"], probs, vot["idxs"], vot["dims"], vot["corner"], float(vot["res"]),
                        72, adaptive)
    ref = vot[f"grid_adaptive{int(adaptive)}"]
    np.testing.assert_allclose(g, ref, rtol=1e-4, atol=1e-3)
    assert int(np.argmax(g)) == int(vot[f"argmax_adaptive{int(adaptive)}"])
    assert clib.grid_argmax(g) == int(np.argmax(g))
    g64 = clib.ppf_voting(vot["pc"], vot["tr"], probs, vot["idxs"], vot["dims"], vot["corner"], float(vot["res"]),
                          72, adaptive, f64=True)
    np.testing.assert_allclose(g64, ref, rtol=1e-4, atol=1e-3)


def test_vote_oracle_random_bins_and_probs(vot):
    g = clib.ppf_voting(vot["pc"], vot["tr_rand"], vot["probs_rand"], vot["idxs"], vot["dims"], vot["corner"],
                        float(vot["res"]), 72, True)
    np.testing.assert_allclose(g, vot["grid_rand"], rtol=1e-4, atol=1e-4)


def test_backvote_oracle_matches_reference_kernel(vot):
    res = float(vot["res"])
    out = clib.backvote(vot["pc"], vot["tr"], vot["idxs"], vot["dims"], vot["corner"], res, vot["centre"], 3 * res)
    ref = vot["backvote"]
    same_mask = np.any(out != 0, -1) == np.any(ref != 0, -1)
    # plain-C port vs the reference string built for the CPU: same libm, contraction may differ -> counted and proven
    assert_masks_equal_up_to_rounding(np.any(out != 0, -1), np.any(ref != 0, -1), vot["pc"], vot["tr"], vot["idxs"], vot["dims"],
                                      vot["corner"], res, vot["centre"], np.float32(3 * res), "C oracle vs CPU-built reference", 4)
    np.testing.assert_allclose(out[same_mask], ref[same_mask], rtol=1e-4, atol=1e-6)
    assert not np.any(out[:8])               # degenerate pairs never written (voting.py:87)


def test_rot_oracle_matches_reference_kernel(vot):
    out = clib.rot_voting(vot["pc"], vot["rot"], vot["idxs"][:96])
    np.testing.assert_allclose(out, vot["rot_candidates"], rtol=1e-4, atol=2e-6)
    live = out[8:]
    np.testing.assert_allclose(np.linalg.norm(live, axis=-1), 1.0, atol=1e-5)


@pytest.mark.parametrize("w", [1, 2])
def test_findpeak_oracle_literal_matches_reference_string(vot, w):
    out = clib.findpeak(vot["grid_adaptive1"], w, literal=True)
    np.testing.assert_allclose(out, vot[f"findpeak_w{w}"], rtol=1e-5, atol=1e-4)
    fixed = clib.findpeak(vot["grid_adaptive1"], w, literal=False)
    assert fixed.shape == out.shape and not np.allclose(fixed, out)
    # x = 0 slab: the literal and intended forms coincide there (dropped term is x*gy*gz = 0)
    np.testing.assert_allclose(fixed[0], out[0], rtol=1e-5, atol=1e-4)


def test_sphere_and_targets_match_reference_functions(glue):
    np.testing.assert_array_equal(ref_model.fibonacci_sphere(480), glue["sphere"])
    tr = ref_model.generate_target_tr(glue["pc"].astype(np.float64), glue["target_idx"])
    np.testing.assert_allclose(tr, glue["target_tr"], rtol=1e-6, atol=1e-7)
    rot = ref_model.generate_target_rot(glue["pc"].astype(np.float64), glue["target_idx"], True)
    np.testing.assert_allclose(rot, glue["target_rot"][:, 0], rtol=1e-5, atol=1e-6)


def test_multinomial_race_equivalence(glue):
    draws = ref_model.sample_bins_race(torch.from_numpy(glue["mn_logits"]), torch.from_numpy(glue["mn_q"]))
    np.testing.assert_array_equal(draws.numpy(), glue["mn_draws"])


def test_cdf_sampler_distribution():
    g = torch.Generator().manual_seed(0)
    logits = torch.tensor([[0.0, 1.0, 2.0, -1.0]]).repeat(40000, 1)
    u = torch.rand(40000, generator=g)
    b = ref_model.sample_bins_cdf(logits, u)
    freq = torch.bincount(b, minlength=4).float() / 40000
    np.testing.assert_allclose(freq.numpy(), torch.softmax(logits[0], -1).numpy(), atol=0.01)


def test_sphere_count_oracle():
    rng = np.random.default_rng(0)
    cand = rng.normal(size=(500, 3)).astype(np.float32)
    cand /= np.linalg.norm(cand, axis=-1, keepdims=True)
    sph = ref_model.fibonacci_sphere(480).astype(np.float32)
    thr = np.float32(np.cos(1.5 / 180 * np.pi))
    counts = clib.sphere_count(cand, sph, thr)
    expect = ((cand @ sph.T) > thr).sum(0)
    assert np.abs(counts - expect).sum() <= 2


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_restatement_against_live_reference_modules(enc):
    import sys
    sys.path.insert(0, "/root/reference")
    from models.model import PPFEncoder
    m = PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).eval()
    m.load_state_dict(split_state(enc, "ppf/"))
    rng = np.random.default_rng(3)
    pc = torch.from_numpy(rng.normal(size=(64, 3)).astype(np.float32) * 0.1)
    nrm = torch.nn.functional.normalize(torch.from_numpy(rng.normal(size=(64, 3)).astype(np.float32)), dim=-1)
    feat = torch.from_numpy(rng.normal(size=(64, 40)).astype(np.float32))
    idxs = rng.integers(0, 64, (300, 2))
    with torch.no_grad():
        ref = m(pc[None], nrm[None], feat[None], idxs=idxs)[0]
    out = ref_model.ppf_encode_idx(pc, nrm, feat, idxs, m.state_dict())
    np.testing.assert_allclose(out.numpy(), ref.numpy(), rtol=1e-5, atol=2e-6)
