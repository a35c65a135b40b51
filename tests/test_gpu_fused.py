"""GPU parity of the fused per-object path (encode+sample, privatised vote, back-vote from
bins, fused orientation histogram, survivor statistics) against the oracle and against the
unfused kernels, which test_gpu_parity.py pins to the reference."""
import numpy as np
import pytest
import torch

from cppf_b200 import fast, model, synth, voting
from cppf_b200.pipeline import PoseConfig, PoseEstimator
from oracle import clib, philox, ref_model
from parity_util import assert_masks_equal_up_to_rounding

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _t(a, dt=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV, dt)


def _setup(n, seed, out_dim=141):
    torch.manual_seed(seed)
    m = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=out_dim).to(DEV).eval()
    with torch.no_grad():
        m.final.weight.mul_(6.0)                     # sharper heads than default init: exercises the CDF search
    pc, nrm = synth.synth_bottle(n, seed)
    feat = torch.randn(n, 40, generator=torch.Generator().manual_seed(seed))
    return m, pc, nrm, feat


@pytest.mark.parametrize("impl", ["tc", "simt"])
@pytest.mark.parametrize("n,p,seed", [(512, 50000, 0), (777, 4099, 1)])
def test_encode_sample_matches_oracle(n, p, seed, impl):
    m, pc, nrm, feat = _setup(n, seed)
    idxs = synth.sample_pairs(n, p, seed).astype(np.int32)
    u = torch.rand(p, 4, generator=torch.Generator().manual_seed(seed + 9))
    with torch.no_grad():
        table = _table(m, feat, impl)
        bins, tail = fast.encode_sample(m, _t(pc), _t(nrm), table, _t(idxs, torch.int32), heads=15, uniforms=u.to(DEV),
                                        impl=impl)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    logits = ref_model.ppf_encode_idx(torch.from_numpy(pc), torch.from_numpy(nrm), feat, idxs, sd)
    bins = bins.cpu().long()
    for h, (c0, nb) in enumerate([(0, 32), (32, 32), (64, 36), (100, 36)]):
        ref = ref_model.sample_bins_cdf(logits[:, c0:c0 + nb], u[:, h], exp2=True)
        mism = int((bins[:, h] != ref).sum())
        assert mism <= max(2, p // 5000), f"head {h}: {mism} of {p} draws differ"     # fp32 ulp at a CDF edge
        assert int((bins[:, h] - ref).abs().max()) <= 1
    np.testing.assert_allclose(tail.cpu().numpy().T, logits[:, 136:].numpy(), rtol=1e-4, atol=2e-5)


def test_encode_tc_last_preactivation_matches_oracle():
    """t = [fc1_2(x2) ; fc0_2(x2) + fc2_2.b] (models/model.py:27-28 of the third ResLayer) from the tcgen05 chain
    -- three composed 3xTF32 MMA steps, no sampling involved -- vs the oracle's layer-by-layer fp32 stack,
    incl. a ragged last tile (p % 128 != 0)."""
    n, p = 300, 128 * 37 + 5
    m, pc, nrm, feat = _setup(n, 3)
    idxs = synth.sample_pairs(n, p, 3).astype(np.int64)
    t_dev = torch.full((p, 32), float("nan"), device=DEV)
    with torch.no_grad():
        fast.encode_sample(m, _t(pc), _t(nrm), _table(m, feat, "tc"), _t(idxs, torch.int64), heads=15, seed=5, impl="tc",
                           dbg_t=t_dev)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    tpc, tn = torch.from_numpy(pc), torch.from_numpy(nrm)
    ia, ib = torch.from_numpy(idxs[:, 0]), torch.from_numpy(idxs[:, 1])
    d = tpc[ia] - tpc[ib]
    dn = torch.norm(d, dim=-1)
    dh = d / (dn[:, None] + 1e-7)
    x = torch.cat([feat[ia], feat[ib], (tn[ia] * dh).sum(-1, keepdim=True), (tn[ib] * dh).sum(-1, keepdim=True),
                   (tn[ia] * tn[ib]).sum(-1, keepdim=True), dn[:, None]], -1)
    for i in range(2):
        x = ref_model.res_layer(x, sd, f"res_layers.{i}")
    F = torch.nn.functional
    want = torch.cat([F.linear(x, sd["res_layers.2.fc1.weight"], sd["res_layers.2.fc1.bias"]),
                      F.linear(x, sd["res_layers.2.fc0.weight"], sd["res_layers.2.fc0.bias"] + sd["res_layers.2.fc2.bias"])], -1)
    np.testing.assert_allclose(t_dev.cpu().numpy(), want.numpy(), rtol=1e-4, atol=3e-5)


@pytest.mark.parametrize("impl", ["tc", "simt"])
def test_encode_sample_dense_philox_and_head_mask(impl):
    n = 96
    m, pc, nrm, feat = _setup(n, 4)
    with torch.no_grad():
        table = _table(m, feat, impl)
        b_seed, t_seed = fast.encode_sample(m, _t(pc), _t(nrm), table, None, heads=15, seed=1234567890123, impl=impl)
        u = philox.pair_uniforms(1234567890123, n * n)
        b_inj, t_inj = fast.encode_sample(m, _t(pc), _t(nrm), table, None, heads=15, uniforms=_t(u), impl=impl)
        b_idx, _ = fast.encode_sample(m, _t(pc), _t(nrm), table, _t(synth.dense_pairs(n), torch.int64), heads=15,
                                      uniforms=_t(u), impl=impl)
        b_tr, none_tail = fast.encode_sample(m, _t(pc), _t(nrm), table, None, heads=fast.HEAD_TR, uniforms=_t(u), impl=impl)
    assert torch.equal(b_seed, b_inj) and torch.equal(t_seed, t_inj)       # kernel Philox == oracle Philox
    assert torch.equal(b_idx, b_inj)                                        # dense enumeration == explicit list
    assert none_tail is None and torch.equal(b_tr[:, :2], b_inj[:, :2]) and not b_tr[:, 2:].any()


@pytest.mark.parametrize("n", [96, 128, 300, 1024, 4096])
def test_dense_tc_encoder_matches_indexed_and_oracle_at_every_tile_shape(n):
    """The BENCHMARKED instantiation -- encode_sample_tc_kernel<dense> with row-aligned 128-pair tiles, the a-side table row
    through the ones-operand MMA, tiles_per_row = ceil(n / 128) (1 ragged tile at 96, 1 full at 128, 2 full + 1 ragged at
    300, 8 at 1024, 32 at the headline N = 4096) -- against (i) the indexed instantiation on the same pairs and uniforms and
    (ii) the oracle's fp32 layer-by-layer MLP + sequential inverse-CDF draw, on a random subset of <= 200 000 pairs.
    Bars: bins equal except at a CDF edge within fp32 rounding (counted, off by at most one bin); tail logits rtol 1e-4."""
    seed = 987654321
    m, pc, nrm, feat = _setup(n, 11)
    with torch.no_grad():
        table = _table(m, feat, "tc")
        b_dense, t_dense = fast.encode_sample(m, _t(pc), _t(nrm), table, None, heads=15, seed=seed, impl="tc")
    n_pairs = n * n
    rng = np.random.default_rng(n)
    sub = np.sort(rng.choice(n_pairs, size=min(n_pairs, 200_000), replace=False)).astype(np.int64)
    # always include the corners of the tiling: first / last pair of a row, tile boundaries, the diagonal, the last row
    special = np.array([0, n - 1, n, n_pairs - 1, n_pairs - n, (n // 2) * n + n // 2] +
                       [r * n + c for r in (0, n // 3, n - 1) for c in (127, 128, 129, n - 2) if 0 <= c < n], np.int64)
    sub = np.unique(np.concatenate([sub, special]))
    idx = np.stack([sub // n, sub % n], -1)
    u = philox.pair_uniforms_at(seed, sub)
    with torch.no_grad():
        b_idx, t_idx = fast.encode_sample(m, _t(pc), _t(nrm), table, _t(idx, torch.int64), heads=15, uniforms=_t(u), impl="tc")
    bd = b_dense.cpu().numpy()[sub].astype(np.int64)
    bi = b_idx.cpu().numpy().astype(np.int64)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    logits = ref_model.ppf_encode_idx(torch.from_numpy(pc), torch.from_numpy(nrm), feat, idx, sd)
    tu = torch.from_numpy(u)
    allowed = max(2, len(sub) // 5000)
    for h, (c0, nb) in enumerate([(0, 32), (32, 32), (64, 36), (100, 36)]):
        ref = ref_model.sample_bins_cdf(logits[:, c0:c0 + nb], tu[:, h], exp2=True).numpy()
        for name, other in (("indexed kernel", bi[:, h]), ("oracle", ref)):
            d = bd[:, h] - other
            mism = int((d != 0).sum())
            print(f"[dense tc n={n}] head {h} vs {name}: {mism} of {len(sub)} draws differ")
            assert mism <= allowed, f"n={n} head {h}: {mism} of {len(sub)} dense draws differ from the {name}"
            assert int(np.abs(d).max()) <= 1
    tail_d = t_dense.cpu().numpy()[:, sub].T
    np.testing.assert_allclose(tail_d, logits[:, 136:].numpy(), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(tail_d, t_idx.cpu().numpy().T, rtol=1e-5, atol=1e-5)     # a-side row via MMA vs FADD: 3xTF32 grade


def test_backvote_bins_bit_equal_to_reference_kernel_dense_1024():
    """All 1 048 576 ordered pairs of an N = 1024 bottle under the trained-like vote load: the survivor mask of the
    bins-driven back-vote kernel (dense enumeration, arc window, fast accept) is BIT-EQUAL to the mask the reference's own
    `backvote` kernel (models/voting.py:74-112, built for sm_100a from the reference string) leaves on the same GPU."""
    from oracle import ref_gpu
    if not ref_gpu.available():
        pytest.skip("reference cubins / cuda-python not present")
    cfg, pc, idxs, tr, corner, dims = _vote_case(1024, 0, 8, dense=True)
    lut = fast.decode_lut(cfg["vote_range"])
    b_mu = torch.argmin((torch.from_numpy(tr[:, 0:1]) - lut[None, :32]).abs(), -1)
    b_nu = torch.argmin((torch.from_numpy(tr[:, 1:2]) - lut[None, 32:64]).abs(), -1)
    bins = torch.stack([b_mu, b_nu, b_mu * 0, b_mu * 0], -1).to(torch.uint8).to(DEV)
    grid = torch.zeros(dims, device=DEV)
    fast.vote_fast(_t(pc), None, grid, _t(corner), cfg["res"], bins=bins, lut=lut.to(DEV))
    flat = voting.grid_argmax(grid)
    gyz = dims[1] * dims[2]
    # the winning cell (nearly every trained-like pair survives) and four cells away from it (selective masks)
    for shift in (0, 4 * gyz + 3 * dims[2] + 2, -(3 * gyz) - 5 * dims[2], 2 * gyz - 7 * dims[2] + 1, 5):
        f2 = (flat + shift).clamp(0, dims[0] * gyz - 1)
        mask = fast.backvote_bins(_t(pc), bins, lut.to(DEV), None, dims, _t(corner), f2, cfg["res"], 3 * cfg["res"])
        cell = np.array(np.unravel_index(int(f2.item()), dims))
        centre = (np.asarray(corner, np.float64) + cell * cfg["res"]).astype(np.float32)          # nocs/inference.py:209,226
        ref_off = ref_gpu.backvote(_t(pc), _t(tr), torch.zeros(idxs.shape[0], 3, device=DEV), _t(idxs, torch.int32), _t(corner),
                                   cfg["res"], 72, dims, _t(centre), 3 * cfg["res"])
        ref_mask = (ref_off != 0).any(-1)
        n_diff = int((mask.bool() != ref_mask).sum().item())
        print(f"[backvote_bins dense 1024 vs reference cubin, cell {cell.tolist()}] survivors {int(ref_mask.sum())} of "
              f"{idxs.shape[0]}, differing bits: {n_diff}")
        assert int(ref_mask.sum()) > 1000
        assert n_diff == 0


def _table(m, feat, impl):
    return m.tc_preproject(feat.to(DEV)) if impl == "tc" else m.preproject(feat.to(DEV))


def _vote_case(n, p, seed, dense=False):
    cfg = synth.BOTTLE
    pc, _ = synth.synth_bottle(n, seed)
    idxs = synth.dense_pairs(n) if dense else synth.sample_pairs(n, p, seed)
    tr = synth.trained_like_tr(pc, idxs)
    corner, dims = synth.vote_grid_geometry(pc, cfg["res"])
    return cfg, pc, idxs, tr, corner, dims


@pytest.mark.parametrize("adaptive", [True, False])
def test_vote_fast_matches_oracle_and_is_deterministic(adaptive):
    cfg, pc, idxs, tr, corner, dims = _vote_case(1024, 200000, 3)
    ref = clib.ppf_voting(pc, tr, np.ones(len(pc), np.float32), idxs.astype(np.int32), dims, corner, cfg["res"], 72,
                          adaptive, f64=True)
    grids = []
    for _ in range(2):
        g = torch.zeros(dims, device=DEV)
        fast.vote_fast(_t(pc), _t(idxs, torch.int32), g, _t(corner), cfg["res"], mu_nu=_t(tr), adaptive=adaptive)
        grids.append(g)
    assert torch.equal(grids[0], grids[1])                                  # integer accumulation: run-to-run identical
    got = grids[0].cpu().numpy()
    # weights are rounded to 2^-14 per corner: |err| <= 2^-15 * contributions.  Against the float64 C oracle a candidate
    # whose grid coordinate sits within float32 rounding of an acceptance bound (models/voting.py:36-39) may be accepted
    # by one side only: at most a few cells may differ, each by less than one vote
    bad = ~np.isclose(got, ref, rtol=1e-4, atol=5e-3)
    print(f"[vote_fast adaptive={adaptive} vs float64 C oracle] cells outside tolerance: {int(bad.sum())} of {bad.size}")
    assert bad.sum() <= 8 and (np.abs(got - ref)[bad] < 1.0).all()
    assert int(voting.grid_argmax(grids[0]).item()) == int(np.argmax(ref))
    from oracle import ref_gpu
    if ref_gpu.available():
        # ... and against the reference's own kernel on this GPU (same float32 arithmetic, contraction spelled out in
        # csrc/common.cuh) every cell agrees: the two accept exactly the same candidates
        rg = ref_gpu.ppf_voting(_t(pc), _t(tr), torch.ones(len(pc), device=DEV), _t(idxs, torch.int32), torch.zeros(dims, device=DEV),
                                _t(corner), cfg["res"], 72, adaptive).cpu().numpy()
        np.testing.assert_allclose(got, rg, rtol=1e-4, atol=5e-3)
        assert int(np.argmax(got)) == int(np.argmax(rg))
    slow = torch.zeros(dims, device=DEV)
    voting.ppf_vote(_t(pc), _t(tr), _t(idxs, torch.int32), slow, _t(corner), cfg["res"], 72, adaptive)
    assert int(voting.grid_argmax(slow).item()) == int(np.argmax(ref))
    assert abs(float(got.sum()) - float(ref.sum())) / float(ref.sum()) < 1e-5


def test_vote_fast_from_bins_dense_and_overflow_flush():
    cfg, pc, idxs, tr, corner, dims = _vote_case(600, 0, 5, dense=True)      # 360k pairs
    lut = fast.decode_lut(cfg["vote_range"])
    b_mu = torch.argmin((torch.from_numpy(tr[:, 0:1]) - lut[None, :32]).abs(), -1)
    b_nu = torch.argmin((torch.from_numpy(tr[:, 1:2]) - lut[None, 32:64]).abs(), -1)
    bins = torch.stack([b_mu, b_nu, torch.zeros_like(b_mu), torch.zeros_like(b_mu)], -1).to(torch.uint8)
    mu_nu = torch.stack([lut[b_mu], lut[32 + b_nu]], -1)
    np.testing.assert_array_equal(mu_nu.numpy(), tr)                         # lut reproduces nocs/inference.py:187-188
    g_bins = torch.zeros(dims, device=DEV)
    fast.vote_fast(_t(pc), None, g_bins, _t(corner), cfg["res"], bins=bins.to(DEV), lut=lut.to(DEV))
    g_mn = torch.zeros(dims, device=DEV)
    fast.vote_fast(_t(pc), _t(idxs, torch.int64), g_mn, _t(corner), cfg["res"], mu_nu=mu_nu.to(DEV))
    assert torch.equal(g_bins, g_mn)
    ref = clib.ppf_voting(pc, tr, np.ones(len(pc), np.float32), idxs.astype(np.int32), dims, corner, cfg["res"], 72, True,
                          f64=True)
    np.testing.assert_allclose(g_bins.cpu().numpy(), ref, rtol=1e-4, atol=5e-3)
    assert int(voting.grid_argmax(g_bins).item()) == int(np.argmax(ref))
    # adversarial overflow case: every pair's 72 non-adaptive candidates fall into one cell
    n = 64
    pts = np.zeros((n, 3), np.float32)
    pts[:] = (0.02, 0.02, 0.02)
    pts[:, 0] += np.linspace(0.0, 0.008, n)
    pts[0] = (0, 0, 0); pts[1] = (0.04, 0.04, 0.04)
    p = 3_000_000
    idx = np.stack([np.full(p, 10), np.full(p, 20)], -1).astype(np.int32)
    mn = np.tile(np.array([[0.0, 1e-6]], np.float32), (p, 1))
    cor, dm = synth.vote_grid_geometry(pts, 4e-3)
    g = torch.zeros(dm, device=DEV)
    fast.vote_fast(_t(pts), _t(idx, torch.int32), g, _t(cor), 4e-3, mu_nu=_t(mn), adaptive=False)
    assert abs(float(g.sum().item()) - 72.0 * p) / (72.0 * p) < 1e-4          # 2.2e8 votes, far beyond one u32 cell


def test_backvote_bins_matches_oracle():
    cfg, pc, idxs, tr, corner, dims = _vote_case(512, 30000, 6)
    lut = fast.decode_lut(cfg["vote_range"])
    b_mu = torch.argmin((torch.from_numpy(tr[:, 0:1]) - lut[None, :32]).abs(), -1)
    b_nu = torch.argmin((torch.from_numpy(tr[:, 1:2]) - lut[None, 32:64]).abs(), -1)
    bins = torch.stack([b_mu, b_nu, b_mu * 0, b_mu * 0], -1).to(torch.uint8).to(DEV)
    grid = torch.zeros(dims, device=DEV)
    fast.vote_fast(_t(pc), _t(idxs, torch.int32), grid, _t(corner), cfg["res"], bins=bins, lut=lut.to(DEV))
    flat = voting.grid_argmax(grid)
    mask = fast.backvote_bins(_t(pc), bins, lut.to(DEV), _t(idxs, torch.int32), dims, _t(corner), flat, cfg["res"],
                              3 * cfg["res"])
    _, centre = ref_model.centre_from_grid(grid.cpu().numpy(), corner, cfg["res"])
    ref = clib.backvote(pc, tr, idxs.astype(np.int32), dims, corner, cfg["res"], centre.astype(np.float32), 3 * cfg["res"])
    # the bins-driven kernel (analytic arc window + fast accept) against the plain-C oracle (gcc + libm): bit-equal except
    # for candidates on the tolerance sphere within float32 rounding (counted and proven)
    assert_masks_equal_up_to_rounding(mask.cpu().numpy().astype(bool), np.any(ref != 0, -1), pc, tr, idxs, dims, corner,
                                      cfg["res"], centre.astype(np.float32), np.float32(3 * cfg["res"]),
                                      "backvote_bins vs C oracle", 4)
    # ... and bit-equal to the reference's own kernel on this GPU, where its cubin travelled
    from oracle import ref_gpu
    if ref_gpu.available():
        ref_off = ref_gpu.backvote(_t(pc), _t(tr), torch.zeros(idxs.shape[0], 3, device=DEV), _t(idxs, torch.int32), _t(corner),
                                   cfg["res"], 72, dims, _t(centre.astype(np.float32)), 3 * cfg["res"])
        np.testing.assert_array_equal(mask.bool().cpu().numpy(), (ref_off != 0).any(-1).cpu().numpy())


def test_rot_hist_equals_unfused_and_stats_match_numpy():
    n, p = 400, 6000
    cfg = synth.BOTTLE
    pc, nrm = synth.synth_bottle(n, 7)
    idxs = synth.sample_pairs(n, p, 7).astype(np.int32)
    rng = np.random.default_rng(7)
    bins = torch.from_numpy(rng.integers(0, 32, (p, 4)).astype(np.uint8)).to(DEV)
    lut = fast.decode_lut(cfg["vote_range"]).to(DEV)
    mask = torch.from_numpy((rng.random(p) < 0.6).astype(np.uint8)).to(DEV)
    mask[:5] = 1
    idxs[:5, 1] = idxs[:5, 0]                                                 # degenerate survivors
    kept, cnt, pos = voting.compact_pairs(mask, _t(idxs, torch.int32), n, want_pos=True)
    c = int(cnt.item())
    sphere = torch.from_numpy(ref_model.fibonacci_sphere(480).astype(np.float32)).to(DEV)
    thr = float(np.float32(np.cos(1.5 / 180 * np.pi)))
    for which in (0, 1):
        counts = fast.rot_hist(_t(pc), bins, lut, _t(idxs, torch.int32), pos, cnt, sphere, which=which, max_samples=p, thr=thr)
        rot = lut[(64 if which == 0 else 100) + bins[pos[:c], 2 + which].long()]
        cand = voting.rot_vote(_t(pc), rot.contiguous(), kept[:c].contiguous(), 72)
        ref = voting.sphere_count(cand, sphere, thr)
        assert torch.equal(counts.int(), ref)
        # sub-sample: without replacement, exactly max_samples pairs
        sub = fast.rot_hist(_t(pc), bins, lut, _t(idxs, torch.int32), pos, cnt, sphere, which=which, max_samples=1000,
                            offset_seed=5, thr=thr)
        assert 0 < float(sub.sum()) < float(counts.sum())
    tail = torch.randn(5, p, generator=torch.Generator().manual_seed(1)).to(DEV)
    best_up = torch.tensor([17], device=DEV)
    best_right = torch.tensor([333], device=DEV)
    st = fast.survivor_stats(_t(pc), _t(nrm), tail, _t(idxs, torch.int32), pos, cnt, sphere, best_up, best_right).cpu().numpy()
    sel = pos[:c].cpu().numpy()
    tl = tail.cpu().numpy()
    assert st[3] == c
    np.testing.assert_allclose(st[:3], tl[2:5, sel].sum(1), rtol=1e-6, atol=1e-4)
    sph = sphere.cpu().numpy()
    ab = pc[idxs[sel, 0]] - pc[idxs[sel, 1]]
    abn = ab / (np.sqrt((ab ** 2).sum(-1)) + np.float32(1e-7))[:, None]
    pn = nrm[idxs[sel, 0]].copy()
    pn[(pn * abn).sum(-1) < 0] *= -1
    for k, b in ((4, 17), (5, 333)):
        t = np.where((pn * sph[b]).sum(-1) > 0, 1.0, -1.0)
        np.testing.assert_allclose(st[k], (tl[k - 4, sel] * t).sum(), rtol=1e-5, atol=1e-3)
        # the sign rule equals the reference's BCE comparison (nocs/inference.py:295-302)
        _, up_l, down_l = ref_model.aux_sign(pc, nrm, idxs[sel], sph[b].astype(np.float64), tl[k - 4, sel])
        assert (st[k] < 0) == (down_l < up_l)


@pytest.mark.parametrize("regress_right", [False, True])
def test_fused_pipeline_equals_two_pass_pipeline(regress_right):
    torch.manual_seed(0)
    pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(DEV).eval()
    ppf = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).to(DEV).eval()
    cfg = PoseConfig.from_dict(dict(synth.BOTTLE, n_pairs=30000, rot_subsample=0, regress_right=regress_right))
    est = PoseEstimator(pe, ppf, cfg, DEV)
    n = 700
    pc, nrm = synth.synth_bottle(n, 11)
    idxs = synth.sample_pairs(n, cfg.n_pairs, 11).astype(np.int32)
    u = torch.rand(cfg.n_pairs, 4, generator=torch.Generator().manual_seed(2)).to(DEV)
    fused = est.estimate_fused(pc, nrm, seed=0, idxs=idxs, uniforms=u, return_debug=True)
    # two-pass path with the same draws: survivors' rotation uniforms are the rows of u they came from
    two = est.estimate(pc, nrm, seed=0, idxs=idxs, noise={"u_mu": u[:, 0].contiguous(), "u_nu": u[:, 1].contiguous()},
                       return_debug=True)
    assert fused["argmax_flat"] == int(two["argmax"].item())
    np.testing.assert_allclose(fused["T_host"], two["T_host"], rtol=0, atol=1e-9)
    assert abs(fused["n_survivors"] - two["n_survivors"]) <= 3
    np.testing.assert_allclose(fused["pred_scale"], two["pred_scale"], rtol=1e-3)
    # fused vote grid vs the float-reduction grid of the two-pass path: identical up to the 2^-14 weight rounding,
    # except where one of the 60k draws fell on a CDF edge (exp2f vs expf) and moved a whole circle of votes
    diff = np.abs(fused["grid"].cpu().numpy() - two["grid"].cpu().numpy())
    assert (diff > 5e-3).mean() < 0.05
    assert np.isfinite(fused["RT"]).all() and abs(np.linalg.norm(fused["up"]) - 1) < 1e-6


@pytest.mark.parametrize("regress_right,dense", [(False, False), (True, False), (False, True)])
def test_one_call_pipeline_equals_staged_pipeline(regress_right, dense):
    """cppf_pose_fused (one library call, geometry derived on the device, no host round trip) against the same
    kernels launched stage by stage from Python with host-side geometry: identical record."""
    torch.manual_seed(0)
    pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(DEV).eval()
    ppf = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).to(DEV).eval()
    n = 300 if dense else 700
    n_pairs = 0 if dense else 30000
    cfg = PoseConfig.from_dict(dict(synth.BOTTLE, n_pairs=n_pairs, rot_subsample=3000, regress_right=regress_right))
    est = PoseEstimator(pe, ppf, cfg, DEV)
    pc, nrm = synth.synth_bottle(n, 13)
    idxs = None if dense else synth.sample_pairs(n, n_pairs, 13).astype(np.int32)
    p = n * n if dense else n_pairs
    u = torch.rand(p, 4, generator=torch.Generator().manual_seed(3)).to(DEV)
    inj = None
    if dense:       # trained-like bins, as bench.py injects them
        inj = synth.trained_like_bins_dense_torch(torch.from_numpy(pc).to(DEV), synth.BOTTLE)
    staged = est.estimate_fused(pc, nrm, seed=4, idxs=idxs, uniforms=u, inject_bins=inj, staged=True)
    for src in ((pc, nrm), (torch.from_numpy(pc).to(DEV), torch.from_numpy(nrm).to(DEV))):     # host and resident inputs
        one = est.estimate_fused(src[0], src[1], seed=4, idxs=idxs, uniforms=u, inject_bins=inj)
        assert one["argmax_flat"] == staged["argmax_flat"]
        assert one["best_bins"] == staged["best_bins"]
        assert one["n_survivors"] == staged["n_survivors"] > 0
        np.testing.assert_allclose(one["T_host"], staged["T_host"], rtol=0, atol=0)
        np.testing.assert_allclose(one["pred_scale"], staged["pred_scale"], rtol=1e-6)
        np.testing.assert_allclose(one["RT"], staged["RT"], rtol=1e-6, atol=1e-9)     # scale sums: different summation order
    # asynchronous use: several objects in flight, records read afterwards
    pend = [est.estimate_fused(pc, nrm, seed=4, idxs=idxs, uniforms=u, inject_bins=inj, sync=False) for _ in range(3)]
    for q in pend:
        assert q.result()["argmax_flat"] == staged["argmax_flat"]


def test_one_call_pipeline_reports_oversized_grid():
    torch.manual_seed(0)
    pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(DEV).eval()
    ppf = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).to(DEV).eval()
    est = PoseEstimator(pe, ppf, PoseConfig.from_dict(dict(synth.BOTTLE, n_pairs=2000)), DEV)
    pc, nrm = synth.synth_bottle(256, 1)
    with pytest.raises(RuntimeError, match="capacity"):
        est.estimate_fused(torch.from_numpy(pc).to(DEV), torch.from_numpy(nrm).to(DEV), max_cells=1000)


def test_vote_routed_64_cube_matches_oracle_and_global_kernel():
    """64^3 grid (BASELINE config 3): too large for one SM's shared memory -> candidates are routed through
    HBM to 7 x-slabs (cppf_vote_routed).  Same fixed-point sums as the privatised kernel: exact vs the oracle
    up to the 2^-14 weight rounding, deterministic, same argmax as the global-reduction kernel."""
    res = 4e-3
    n, p = 1024, 120000
    pc, _ = synth.synth_cylinder_grid64(n, 2, res=res)
    corner, dims = synth.vote_grid_geometry(pc, res)
    assert dims == (64, 64, 64) and not fast.vote_fits_private(dims) and fast.vote_routed_supported(dims)
    idxs = synth.sample_pairs(n, p, 2)
    tr = synth.trained_like_tr(pc, idxs)
    ref = clib.ppf_voting(pc, tr, np.ones(n, np.float32), idxs.astype(np.int32), dims, corner, res, 72, True, f64=True)
    grids = []
    for _ in range(2):
        g = torch.zeros(dims, device=DEV)
        fast.vote_routed(_t(pc), _t(idxs, torch.int32), g, _t(corner), res, mu_nu=_t(tr))
        grids.append(g)
    assert torch.equal(grids[0], grids[1])
    got = grids[0].cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=5e-3)
    assert int(voting.grid_argmax(grids[0]).item()) == int(np.argmax(ref))
    assert abs(float(got.sum()) - float(ref.sum())) / float(ref.sum()) < 1e-5
    # dense pairs from bins, two passes over a deliberately small pool (scratch sized for 60 % of the pairs)
    n2 = 640
    pc2, _ = synth.synth_cylinder_grid64(n2, 3, res=res)
    corner2, dims2 = synth.vote_grid_geometry(pc2, res)
    inj = synth.trained_like_bins_dense_torch(torch.from_numpy(pc2).to(DEV), dict(synth.BOTTLE))
    bins = torch.zeros(n2 * n2, 4, dtype=torch.uint8, device=DEV)
    bins[:, :3] = inj
    lut = fast.decode_lut(synth.BOTTLE["vote_range"]).to(DEV)
    full = torch.zeros(dims2, device=DEV)
    fast.vote_routed(_t(pc2), None, full, _t(corner2), res, bins=bins, lut=lut)
    from cppf_b200 import _lib
    nb = _lib.lib().cppf_vote_routed_scratch_bytes(n2 * n2 * 6 // 10, 72, *dims2)
    small = torch.zeros(dims2, device=DEV)
    fast.vote_routed(_t(pc2), None, small, _t(corner2), res, bins=bins, lut=lut,
                     scratch=torch.empty(nb, dtype=torch.uint8, device=DEV))
    assert torch.equal(full, small)
    # slab passes (the privatised kernel with one x-slab per CTA): bit-identical to the routed variant
    slabs = torch.zeros(dims2, device=DEV)
    fast.vote_slabs(_t(pc2), None, slabs, _t(corner2), res, bins=bins, lut=lut)
    assert torch.equal(full, slabs)
    b = bins.long()
    mu_nu = torch.stack([lut[b[:, 0]], lut[32 + b[:, 1]]], -1).contiguous()
    slow = torch.zeros(dims2, device=DEV)
    voting.ppf_vote(_t(pc2), mu_nu, None, slow, _t(corner2), res, 72, True)
    np.testing.assert_allclose(full.cpu().numpy(), slow.cpu().numpy(), rtol=2e-4, atol=2e-2)
    assert int(voting.grid_argmax(full).item()) == int(voting.grid_argmax(slow).item())
