"""The back-vote kernel (csrc/vote_private.cu, backvote_bins_kernel) does not walk all n candidates of a pair like
models/voting.py:99-110 does: it bounds analytically -- with approximate square roots / divisions -- the arc of the
circle that can come within `tol` of the winning centre, pre-filters candidates with two FMAs, and accepts clear hits
early.  Each shortcut only has to be CONSERVATIVE (the candidates it keeps are still tested with the reference's own
arithmetic).  This CPU test restates the shortcuts in float32 numpy, perturbs the approximate operations by a few ulp,
and checks on many random pairs and centres that (1) every candidate the reference test accepts lies inside the
window, (2) passes the pre-filter, and (3) every early accept is a true hit."""
import numpy as np
import pytest

from cppf_b200 import synth

F = np.float32


def _frames(pc, idxs):
    a, b = pc[idxs[:, 0]], pc[idxs[:, 1]]
    ab = (a - b).astype(F)
    ln = np.sqrt((ab * ab).sum(-1, dtype=F)).astype(F)
    ok = ln.astype(np.float64) >= 1e-7                                     # models/voting.py:84
    ab = (ab / (ln.astype(np.float64) + 1e-7).astype(F)[:, None]).astype(F)
    co = np.stack([np.zeros_like(ln), -ab[:, 2], ab[:, 1]], -1).astype(F)
    lc = np.sqrt((co * co).sum(-1, dtype=F)).astype(F)
    alt = lc.astype(np.float64) < 1e-7
    co[alt] = np.stack([-ab[alt, 1], ab[alt, 0], np.zeros(alt.sum(), F)], -1)
    lc = np.sqrt((co * co).sum(-1, dtype=F)).astype(F)
    ex = (co / (lc.astype(np.float64) + 1e-7).astype(F)[:, None]).astype(F)
    return a.astype(F), ab, ex, ok


def _jitter(x, rng, ulps=4):
    """x (1 + e), |e| <= ulps * 2^-24: stands in for sqrt.approx / __fdividef."""
    return (x * (1 + rng.uniform(-ulps, ulps, x.shape) * 2.0 ** -24)).astype(F)


@pytest.mark.parametrize("votes", ["trained_like", "random"])
def test_window_prefilter_and_early_accept_are_conservative(votes):
    rng = np.random.default_rng(7)
    n_pts, p = 2048, 150_000
    res = F(synth.BOTTLE["res"])
    tol = F(3) * res                                                        # nocs/inference.py:226
    pc, _ = synth.synth_bottle(n_pts, 3)
    idxs = synth.sample_pairs(n_pts, p, 4)
    if votes == "trained_like":
        mu_nu = synth.trained_like_tr(pc, idxs)
    else:
        t = np.arange(32, dtype=F)
        mu_nu = np.stack([(t / F(31) * F(0.5) - F(0.25))[rng.integers(0, 32, p)], (t / F(31) * F(0.25))[rng.integers(0, 32, p)]], -1)
    mu, nu = mu_nu[:, 0].astype(F), mu_nu[:, 1].astype(F)
    a, ab, ex, ok = _frames(pc, idxs)
    c = (a - ab * mu[:, None]).astype(F)
    x = (ex * nu[:, None]).astype(F)
    y = np.cross(x, ab).astype(F)
    ey = np.cross(ex, ab).astype(F)
    n = np.minimum(((nu / res).astype(np.float64) * (2 * np.pi)).astype(np.int64), 72)      # models/voting.py:97
    n = np.maximum(n, 0)
    # centres: the object's centre cell (where trained-like votes concentrate) and cells scattered around it
    corner, dims = synth.vote_grid_geometry(pc, float(res))
    centre_cell = np.floor((0.0 - corner) / res).astype(int)
    checked_hits = early = 0
    for trial in range(4):
        cell = centre_cell + (rng.integers(-3, 4, 3) if trial else 0)
        T = (corner.astype(np.float64) + cell * float(res)).astype(F)
        v = (T[None] - c).astype(F)
        dpl, px, py = (v * ab).sum(-1, dtype=F), (v * ex).sum(-1, dtype=F), (v * ey).sum(-1, dtype=F)
        vv = (v * v).sum(-1, dtype=F)
        tol2 = tol * tol
        kq = (F(0.5) * (nu * nu + vv - tol2) - (F(0.02) * tol2 + F(1e-6) * (nu * nu + vv))).astype(F)
        unit = (ex * ex).sum(-1, dtype=F) > F(0.9999)
        # ---- window (kernel: `if (n > 12 && unit_frame)`)
        i_lo = np.zeros(p, np.int64)
        i_cnt = n.copy()
        rho = _jitter(np.sqrt(px * px + py * py), rng)
        room = (tol2 * F(1.01) + F(1e-12) - dpl * dpl - (nu - rho) * (nu - rho)).astype(F)
        two_nr = (F(2) * np.abs(nu) * rho).astype(F)
        use = (n > 12) & unit
        none = use & (room < 0)
        i_cnt[none] = 0
        arc = use & ~none & (room < F(1.9) * two_nr)
        with np.errstate(divide="ignore", invalid="ignore"):
            q = _jitter(room / two_nr, rng)
            dmax = (_jitter(np.sqrt(F(2) * q), rng) * (F(0.22) * q + F(1)) + F(1e-5)).astype(F)
            inv_step = (n.astype(F) * F(0.159154943)).astype(F)
            apx, apy = np.abs(px), np.abs(py)
            mx, mn = np.maximum(apx, apy), np.minimum(apx, apy)
            t = _jitter(mn / np.maximum(mx, F(1e-30)), rng)
            t2 = t * t
            th = (((F(-0.0464964749) * t2 + F(0.15931422)) * t2 + F(-0.327622764)) * t2 * t + t).astype(F)
            th = np.where(apy > apx, F(1.57079633) - th, th)
            th = np.where(px < 0, F(3.14159265) - th, th)
            th = np.where(py < 0, -th, th)
            th = np.where(nu < 0, th + F(3.14159265), th)
            th = np.where(th < 0, th + F(6.2831853), th).astype(F)
            w = (dmax * inv_step + F(0.52)).astype(np.int64) + 1
            lo = (th * inv_step + F(0.5)).astype(np.int64) - w
        narrow = arc & (2 * w + 1 < n)
        i_lo[narrow] = lo[narrow]
        i_cnt[narrow] = 2 * w[narrow] + 1
        assert np.all(i_lo[narrow] >= -n[narrow] // 2 - 1) and np.all(i_lo[narrow] <= n[narrow])   # one wrap suffices
        marg = int(tol / res) + 2
        interior = bool(np.all(cell >= marg) and np.all(cell < np.array(dims) - 1 - marg))
        # ---- the reference's candidates (models/voting.py:99-104), i = 0 .. n-1
        for i in range(72):
            live = ok & (i < n)
            if not live.any():
                break
            ang = ((i * 2) * np.pi / np.maximum(n, 1).astype(np.float64)).astype(F)
            cs, sn = np.cos(ang).astype(F), np.sin(ang).astype(F)
            cand = (c + x * cs[:, None] + y * sn[:, None]).astype(F)
            d = np.sqrt(((cand - T[None]) ** 2).sum(-1, dtype=F))
            hit = live & (d <= tol)
            rel = (i - i_lo) % np.maximum(n, 1)
            in_window = rel < i_cnt
            bad = hit & ~in_window
            assert not bad.any(), f"{int(bad.sum())} hits outside the window (trial {trial}, i={i})"
            qv = (nu * (px * cs + py * sn)).astype(F)
            assert not (hit & unit & (qv < kq)).any(), "pre-filter rejects a hit"
            if interior:
                acc = live & unit & in_window & (qv >= kq) & (nu != 0) & ((nu * nu + vv - F(2) * qv) <= F(0.9) * tol2)
                assert not (acc & ~(d <= tol)).any(), "early accept of a candidate farther than tol"
                early += int(acc.sum())
            checked_hits += int(hit.sum())
    assert checked_hits > 1000                                               # the cases are not vacuous
    if votes == "trained_like":
        assert early > 1000
