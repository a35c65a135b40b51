"""GPU parity: every CUDA entry point (through the C ABI / Python mirror) against the
oracle on the same seeded inputs, against the golden fixtures minted from the reference,
and -- where the prebuilt cubins travelled -- against the reference's own kernels running
on the same GPU.  Bars: argmax / masks / bins / counts bit-exact; fp32 results within the
tolerance written at each assert (north star: 1e-4 relative)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, split_state
from cppf_b200 import model, synth, voting
from oracle import clib, ref_model
from parity_util import assert_masks_equal_up_to_rounding

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _t(a, dt=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV, dt)


@pytest.fixture(scope="module")
def enc():
    d = load_golden("encoder_bottle.npz")
    m = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).to(DEV).eval()
    m.load_state_dict(split_state(d, "ppf/"))
    return d, m


@pytest.fixture(scope="module")
def vot():
    return load_golden("voting_bottle.npz")


# ------------------------------------------------------------------ pair MLP
def test_encode_idx_matches_reference_fixture(enc):
    d, m = enc
    with torch.no_grad():
        out = m(_t(d["pc"])[None], _t(d["nrm"])[None], _t(d["feat"])[None], idxs=d["idxs"])
    assert out.shape == (1, 512, 141)
    np.testing.assert_allclose(out[0].cpu().numpy(), d["logits"], rtol=1e-4, atol=1e-5)


def test_encode_dense_matches_reference_fixture(enc):
    d, m = enc
    n = int(d["dense_n"])
    with torch.no_grad():
        out = m(_t(d["pc"][:n])[None], _t(d["nrm"][:n])[None], _t(d["feat"][:n])[None], dist=_t(d["dense_dist"])[None])
        out_exact = m(_t(d["pc"][:n])[None], _t(d["nrm"][:n])[None], _t(d["feat"][:n])[None])
    assert out.shape == (1, n, n, 141)
    np.testing.assert_allclose(out[0].cpu().numpy(), d["dense_logits"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out_exact[0].cpu().numpy(), d["dense_logits"], rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("n,p,seed", [(1024, 100000, 0), (4096, 33333, 1), (97, 31, 2), (64, 1, 3)])
def test_encode_idx_matches_oracle(n, p, seed):
    torch.manual_seed(seed)
    m = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).to(DEV).eval()
    pc, nrm = synth.synth_bottle(n, seed)
    feat = torch.randn(n, 40, generator=torch.Generator().manual_seed(seed)).numpy()
    idxs = synth.sample_pairs(n, p, seed)
    with torch.no_grad():
        out = m.forward_with_idx(_t(pc), _t(nrm), _t(feat), idxs)
        out32 = m.forward_with_idx(_t(pc), _t(nrm), _t(feat), torch.from_numpy(idxs).to(DEV, torch.int32))
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    ref = ref_model.ppf_encode_idx(torch.from_numpy(pc), torch.from_numpy(nrm), torch.from_numpy(feat), idxs, sd)
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=1e-4, atol=2e-5)
    assert torch.equal(out, out32)


def test_encode_dense_equals_indexed_and_column_window():
    torch.manual_seed(5)
    m = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).to(DEV).eval()
    n = 150
    pc, nrm = synth.synth_bottle(n, 5)
    feat = torch.randn(n, 40, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        dense = m(_t(pc)[None], _t(nrm)[None], feat.to(DEV)[None])[0].reshape(n * n, 141)
        idx = m.forward_with_idx(_t(pc), _t(nrm), feat.to(DEV), synth.dense_pairs(n))
        tr = m._encode(_t(pc), _t(nrm), feat.to(DEV), synth.dense_pairs(n), None, cols=(0, 64))
        up = m._encode(_t(pc), _t(nrm), feat.to(DEV), None, None, cols=(64, 36))
    assert torch.equal(dense, idx)
    assert torch.equal(tr, dense[:, :64])
    assert torch.equal(up, dense[:, 64:100])


def test_encode_other_out_dim_regression_head():
    torch.manual_seed(6)
    m = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=9).to(DEV).eval()      # zero_shot.ipynb head
    n, p = 300, 5000
    pc, nrm = synth.synth_bottle(n, 6)
    feat = torch.randn(n, 40, generator=torch.Generator().manual_seed(6))
    idxs = synth.sample_pairs(n, p, 6)
    with torch.no_grad():
        out = m.forward_with_idx(_t(pc), _t(nrm), feat.to(DEV), idxs)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    ref = ref_model.ppf_encode_idx(torch.from_numpy(pc), torch.from_numpy(nrm), feat, idxs, sd)
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=1e-4, atol=2e-5)


def test_encoder_refuses_grad_and_cpu():
    m = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).to(DEV)
    pc, nrm = synth.synth_bottle(32, 0)
    with pytest.raises(RuntimeError):
        m.forward_with_idx(_t(pc), _t(nrm), torch.zeros(32, 40, device=DEV), synth.sample_pairs(32, 8, 0))
    with torch.no_grad(), pytest.raises(RuntimeError):
        m.forward_with_idx(torch.from_numpy(pc), torch.from_numpy(nrm), torch.zeros(32, 40), synth.sample_pairs(32, 8, 0))


# ------------------------------------------------------------------ centre voting + argmax
def _vote(vot, tr, probs, adaptive, idx_dtype=torch.int32):
    dims = tuple(int(v) for v in vot["dims"])
    grid = torch.zeros(dims, device=DEV)
    voting.ppf_vote(_t(vot["pc"]), _t(tr), _t(vot["idxs"], idx_dtype), grid, _t(vot["corner"]), float(vot["res"]), 72,
                    adaptive, None if probs is None else _t(probs))
    return grid


@pytest.mark.parametrize("adaptive", [True, False])
def test_vote_matches_reference_fixture_and_argmax_is_exact(vot, adaptive):
    grid = _vote(vot, vot["tr"], np.ones(vot["pc"].shape[0], np.float32), adaptive)
    ref = vot[f"grid_adaptive{int(adaptive)}"]
    np.testing.assert_allclose(grid.cpu().numpy(), ref, rtol=1e-4, atol=1e-3)
    flat = voting.grid_argmax(grid)
    assert int(flat.item()) == int(vot[f"argmax_adaptive{int(adaptive)}"]) == int(np.argmax(grid.cpu().numpy()))
    grid64 = _vote(vot, vot["tr"], None, adaptive, torch.int64)      # probs=None == all ones (inference.py:201)
    np.testing.assert_allclose(grid64.cpu().numpy(), ref, rtol=1e-4, atol=1e-3)


def test_vote_random_bins_and_probs(vot):
    grid = _vote(vot, vot["tr_rand"], vot["probs_rand"], True)
    np.testing.assert_allclose(grid.cpu().numpy(), vot["grid_rand"], rtol=1e-4, atol=1e-4)


def test_vote_dense_enumeration_matches_oracle():
    cfg = synth.BOTTLE
    n = 200
    pc, _ = synth.synth_bottle(n, 3)
    idxs = synth.dense_pairs(n)
    tr = synth.trained_like_tr(pc, idxs)
    corner, dims = synth.vote_grid_geometry(pc, cfg["res"])
    grid = torch.zeros(dims, device=DEV)
    voting.ppf_vote(_t(pc), _t(tr), None, grid, _t(corner), cfg["res"], 72, True)
    ref = clib.ppf_voting(pc, tr, np.ones(n, np.float32), idxs.astype(np.int32), dims, corner, cfg["res"], 72, True, f64=True)
    np.testing.assert_allclose(grid.cpu().numpy(), ref, rtol=2e-4, atol=2e-3)
    assert int(voting.grid_argmax(grid).item()) == int(np.argmax(ref))


def test_vote_64cube_grid_and_large_rots_fallback():
    n, p = 2048, 50000
    pc, _ = synth.synth_cylinder_grid64(n, 0)
    corner, dims = synth.vote_grid_geometry(pc, 4e-3)
    assert dims == (64, 64, 64)
    idxs = synth.sample_pairs(n, p, 0).astype(np.int32)
    tr = synth.trained_like_tr(pc, idxs, 32, (0.25, 0.25))
    for n_rots in (72, 90):                                   # 90 > table size -> in-loop angle path
        grid = torch.zeros(dims, device=DEV)
        voting.ppf_vote(_t(pc), _t(tr), _t(idxs, torch.int32), grid, _t(corner), 4e-3, n_rots, True)
        ref = clib.ppf_voting(pc, tr, np.ones(n, np.float32), idxs, dims, corner, 4e-3, n_rots, True, f64=True)
        np.testing.assert_allclose(grid.cpu().numpy(), ref, rtol=2e-4, atol=2e-3)
        assert int(voting.grid_argmax(grid).item()) == int(np.argmax(ref))


def test_argmax_ties_pick_first_index():
    g = torch.zeros(7, 9, 11, device=DEV)
    g.view(-1)[[500, 123, 600]] = 4.0
    flat, val = voting.grid_argmax(g, with_value=True)
    assert int(flat.item()) == 123 and float(val.item()) == 4.0
    g2 = -torch.rand(5, 5, 5, device=DEV) - 1
    assert int(voting.grid_argmax(g2).item()) == int(np.argmax(g2.cpu().numpy()))


def test_vote_against_reference_kernel_on_gpu(vot):
    from oracle import ref_gpu
    if not ref_gpu.available():
        pytest.skip("reference cubins / cuda-python not present")
    dims = tuple(int(v) for v in vot["dims"])
    pts, tr, idx, cor = _t(vot["pc"]), _t(vot["tr"]), _t(vot["idxs"], torch.int32), _t(vot["corner"])
    probs = torch.ones(vot["pc"].shape[0], device=DEV)
    ref = ref_gpu.ppf_voting(pts, tr, probs, idx, torch.zeros(dims, device=DEV), cor, float(vot["res"]), 72, True)
    ours = voting.ppf_vote(pts, tr, idx, torch.zeros(dims, device=DEV), cor, float(vot["res"]), 72, True, probs)
    torch.cuda.synchronize()
    np.testing.assert_allclose(ours.cpu().numpy(), ref.cpu().numpy(), rtol=1e-4, atol=1e-3)
    assert int(torch.argmax(ref).item()) == int(voting.grid_argmax(ours).item())
    # back-vote and orientation candidates against the reference kernels, same device
    centre = _t(vot["centre"])
    ref_off = ref_gpu.backvote(pts, tr, torch.zeros(idx.shape[0], 3, device=DEV), idx, cor, float(vot["res"]), 72, dims,
                               centre, 3 * float(vot["res"]))
    off, mask = voting.backvote(pts, tr, idx, dims, cor, float(vot["res"]), centre, 3 * float(vot["res"]))
    torch.cuda.synchronize()
    # same GPU, same libdevice, and csrc/common.cuh spells out the FMA contraction nvcc chose for the reference string:
    # the survivor mask AND the offsets are bit-equal (P = 100 000 pairs of the fixture)
    np.testing.assert_array_equal(mask.bool().cpu().numpy(), (ref_off != 0).any(-1).cpu().numpy())
    np.testing.assert_array_equal(off.cpu().numpy(), ref_off.cpu().numpy())
    rot = _t(vot["rot"])
    ref_up = ref_gpu.rot_voting(pts, rot, torch.zeros(96, 72, 3, device=DEV), idx[:96].contiguous(), 72)
    up = voting.rot_vote(pts, rot, idx[:96].contiguous(), 72)
    torch.cuda.synchronize()
    np.testing.assert_allclose(up.cpu().numpy(), ref_up.cpu().numpy(), rtol=0, atol=2e-7)
    g = _t(vot["grid_adaptive1"])
    ref_fp = ref_gpu.findpeak(g, torch.zeros_like(g), 1)
    torch.cuda.synchronize()
    np.testing.assert_allclose(voting.findpeak(g, 1, literal=True).cpu().numpy(), ref_fp.cpu().numpy(), rtol=1e-6, atol=1e-5)


# ------------------------------------------------------------------ back-vote, compaction
def test_backvote_matches_reference_fixture_and_compaction(vot):
    res = float(vot["res"])
    dims = tuple(int(v) for v in vot["dims"])
    idx = _t(vot["idxs"], torch.int32)
    off, mask = voting.backvote(_t(vot["pc"]), _t(vot["tr"]), idx, dims, _t(vot["corner"]), res, _t(vot["centre"]), 3 * res)
    ref = vot["backvote"]
    ref_mask = np.any(ref != 0, -1)
    got_mask = mask.cpu().numpy().astype(bool)
    # the fixture was minted from the reference string built for the CPU (gcc + libm; this container has no GPU): bits may
    # differ only for candidates ON the tolerance sphere within float32 rounding -- counted and proven, not waved through
    assert_masks_equal_up_to_rounding(got_mask, ref_mask, vot["pc"], vot["tr"], vot["idxs"], dims, vot["corner"], res,
                                      vot["centre"], np.float32(3 * res), "backvote vs CPU-built reference fixture", 8)
    same = got_mask == ref_mask
    np.testing.assert_allclose(off.cpu().numpy()[same], ref[same], rtol=1e-4, atol=1e-6)
    assert not off[:8].any() and not mask[:8].any()          # degenerate pairs (voting.py:87)
    kept, cnt, pos = voting.compact_pairs(mask, idx, vot["pc"].shape[0], want_pos=True)
    c = int(cnt.item())
    assert c == int(got_mask.sum())
    np.testing.assert_array_equal(kept[:c].cpu().numpy(), vot["idxs"][got_mask])
    np.testing.assert_array_equal(pos[:c].cpu().numpy(), np.nonzero(got_mask)[0])


@pytest.mark.parametrize("n_pairs", [1, 2047, 2048, 2049, 1_000_003])
def test_compaction_sizes_and_dense_enumeration(n_pairs):
    rng = np.random.default_rng(n_pairs)
    mask = (rng.random(n_pairs) < 0.37).astype(np.uint8)
    idxs = rng.integers(0, 1000, (n_pairs, 2)).astype(np.int64)
    kept, cnt, _ = voting.compact_pairs(_t(mask, torch.uint8), _t(idxs, torch.int64), 1000)
    c = int(cnt.item())
    assert c == int(mask.sum())
    np.testing.assert_array_equal(kept[:c].cpu().numpy(), idxs[mask.astype(bool)].astype(np.int32))
    n = 300
    mask = (rng.random(n * n) < 0.1).astype(np.uint8)
    kept, cnt, _ = voting.compact_pairs(_t(mask, torch.uint8), None, n)
    np.testing.assert_array_equal(kept[:int(cnt.item())].cpu().numpy(), synth.dense_pairs(n)[mask.astype(bool)].astype(np.int32))


# ------------------------------------------------------------------ orientation voting
def test_rot_vote_matches_reference_fixture(vot):
    idx = _t(vot["idxs"][:96], torch.int32)
    up = voting.rot_vote(_t(vot["pc"]), _t(vot["rot"]), idx, 72)
    np.testing.assert_allclose(up.cpu().numpy(), vot["rot_candidates"], rtol=1e-4, atol=2e-6)
    assert not up[:8].any()                                   # degenerate pairs untouched (voting.py:130)


def test_sphere_count_matches_oracle_and_mm():
    rng = np.random.default_rng(0)
    cand = rng.normal(size=(72 * 3000 + 17, 3)).astype(np.float32)
    cand /= np.linalg.norm(cand, axis=-1, keepdims=True)
    sph = ref_model.fibonacci_sphere(480).astype(np.float32)
    thr = float(np.float32(np.cos(1.5 / 180 * np.pi)))
    counts = voting.sphere_count(_t(cand), _t(sph), thr).cpu().numpy()
    ref = clib.sphere_count(cand, sph, thr)
    mm = (_t(cand).mm(_t(sph).T) > thr).sum(0).cpu().numpy()            # nocs/inference.py:282-283
    assert np.abs(counts - ref).sum() <= 2 and np.abs(counts - mm).sum() <= 4
    assert int(np.argmax(counts)) == int(np.argmax(ref))


@pytest.mark.parametrize("literal", [True, False])
def test_findpeak_matches_oracle(vot, literal):
    g = vot["grid_adaptive1"]
    for w in (1, 2, 3):
        out = voting.findpeak(_t(g), w, literal=literal).cpu().numpy()
        np.testing.assert_allclose(out, clib.findpeak(g, w, literal=literal), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(voting.findpeak(_t(g), 1, literal=True).cpu().numpy(), vot["findpeak_w1"], rtol=1e-5, atol=1e-4)


# ------------------------------------------------------------------ sampling / decode
def test_sample_race_reproduces_torch_multinomial():
    glue = load_golden("host_glue.npz")
    logits = np.zeros((2000, 141), np.float32)
    logits[:, 32:64] = glue["mn_logits"]
    val, bins = voting.sample_bins(_t(logits), 32, 32, q=_t(glue["mn_q"]), div=31.0, mul_a=0.25, want_bins=True)
    np.testing.assert_array_equal(bins.cpu().numpy(), glue["mn_draws"])
    ref = ref_model.decode_tr(torch.from_numpy(glue["mn_draws"]), torch.from_numpy(glue["mn_draws"]), 32, (0.25, 0.25))[:, 1]
    np.testing.assert_array_equal(val.cpu().numpy(), ref.numpy())


def test_sample_cdf_matches_oracle_and_decode_is_exact():
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(50000, 36, generator=g) * 2
    u = torch.rand(50000, generator=g)
    ref_bins = ref_model.sample_bins_cdf(logits, u)
    vr = 0.7863312261283193                                    # config/category/chair.yaml
    val, bins = voting.sample_bins(logits.to(DEV), 0, 36, u=u.to(DEV), div=31.0, mul_a=2.0, mul_b=vr, sub=vr, want_bins=True)
    mism = (bins.cpu().long() != ref_bins).sum().item()
    assert mism <= 3                                           # expf ulp at a CDF edge
    b = bins.cpu().long()
    expect = b.float() / 31 * 2 * vr - vr                      # nocs/inference.py:187, fp32 step by step
    np.testing.assert_array_equal(val.cpu().numpy(), expect.numpy())
    # Philox stream: deterministic, seed-dependent, right distribution
    l2 = torch.tensor([[0.0, 1.0, 2.0, -1.0]]).repeat(200000, 1).to(DEV)
    _, b1 = voting.sample_bins(l2, 0, 4, seed=7, want_bins=True)
    _, b2 = voting.sample_bins(l2, 0, 4, seed=7, want_bins=True)
    _, b3 = voting.sample_bins(l2, 0, 4, seed=8, want_bins=True)
    assert torch.equal(b1, b2) and not torch.equal(b1, b3)
    freq = torch.bincount(b1.long(), minlength=4).float().cpu() / 200000
    np.testing.assert_allclose(freq.numpy(), torch.softmax(l2[0].cpu(), -1).numpy(), atol=5e-3)


# ------------------------------------------------------------------ drop-in RawKernel call shape
def test_rawkernel_call_shape_with_numpy_and_torch(vot):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin"))
    from models.voting import backvote_kernel, ppf_kernel, rot_voting_kernel      # nocs/inference.py:17
    pc, idxs, tr = vot["pc"], vot["idxs"], vot["tr"]
    dims = tuple(int(v) for v in vot["dims"])
    res = np.float32(vot["res"])
    grid_obj = torch.zeros(dims, device=DEV)
    block_size = (pc.shape[0] ** 2 + 512 - 1) // 512                              # inference.py:192
    ppf_kernel((block_size, 1, 1), (512, 1, 1),
               (_t(pc), _t(tr), _t(np.ones(pc.shape[0], np.float32)), _t(idxs, torch.int32), grid_obj, _t(vot["corner"]),
                res, idxs.shape[0], 72, dims[0], dims[1], dims[2], True))
    np.testing.assert_allclose(grid_obj.cpu().numpy(), vot["grid_adaptive1"], rtol=1e-4, atol=1e-3)
    grid_np = np.zeros(dims, np.float32)                                          # numpy in/out -> copied back
    ppf_kernel((block_size, 1, 1), (512, 1, 1),
               (pc, tr, np.ones(pc.shape[0], np.float32), idxs.astype(np.int32), grid_np, vot["corner"], res,
                idxs.shape[0], 72, dims[0], dims[1], dims[2], True))
    np.testing.assert_allclose(grid_np, vot["grid_adaptive1"], rtol=1e-4, atol=1e-3)
    oc = torch.zeros(idxs.shape[0], 3, device=DEV)
    backvote_kernel(((idxs.shape[0] + 511) // 512, 1, 1), (512, 1, 1),
                    (_t(pc), _t(tr), oc, _t(idxs, torch.int32), _t(vot["corner"]), res, idxs.shape[0], 72, dims[0], dims[1],
                     dims[2], _t(vot["centre"]), np.float32(3 * res)))
    assert_masks_equal_up_to_rounding(np.any(oc.cpu().numpy() != 0, -1), np.any(vot["backvote"] != 0, -1), pc, tr, idxs, dims,
                                      vot["corner"], float(res), vot["centre"], np.float32(3 * res),
                                      "RawKernel-shaped backvote vs CPU-built reference fixture", 8)
    cand = torch.zeros(96, 72, 3, device=DEV)
    rot_voting_kernel((1, 1, 1), (512, 1, 1),
                      (_t(pc), _t(tr), _t(vot["rot"]), cand, _t(idxs[:96], torch.int32), _t(vot["corner"]), res, 96, 72,
                       dims[0], dims[1], dims[2]))
    np.testing.assert_allclose(cand.cpu().numpy(), vot["rot_candidates"], rtol=1e-4, atol=2e-6)
