"""Shared by the parity tests: a float64 restatement of the back-vote candidate test (models/voting.py:78-111) that returns,
per pair, how close the decision is to one of its two thresholds.  Used to PROVE that every mask bit on which two float32
builds of the same expression tree disagree (libm vs libdevice cos/sin, gcc vs nvcc FMA contraction) belongs to a
candidate sitting on the tolerance sphere (or on a grid face) within float32 rounding -- and to count those bits."""
import numpy as np


def backvote_decision_margin(pc, tr, idxs, dims, corner, res, centre, tol, n_rots=72, rows=None):
    """-> float64 [len(rows)]: min over the pair's candidates of the distance (in metres) between the candidate and the
    nearest decision boundary of models/voting.py:102-107 (the tol sphere around the centre, the faces of the grid)."""
    rows = np.arange(len(idxs)) if rows is None else np.asarray(rows)
    out = np.full(len(rows), np.inf)
    pc = pc.astype(np.float64)
    corner = np.asarray(corner, np.float64)
    centre = np.asarray(centre, np.float64)
    hi = (np.asarray(dims, np.float64) - 1.0) * res
    for k, p in enumerate(rows):
        a, b = pc[idxs[p, 0]], pc[idxs[p, 1]]
        ab = a - b
        ln = np.linalg.norm(ab)
        if ln < 1e-7:
            continue
        ab = ab / (ln + 1e-7)
        mu, nu = float(tr[p, 0]), float(tr[p, 1])
        c = a - ab * mu
        co = np.array([0.0, -ab[2], ab[1]])
        if np.linalg.norm(co) < 1e-7:
            co = np.array([-ab[1], ab[0], 0.0])
        x = co / (np.linalg.norm(co) + 1e-7) * nu
        y = np.cross(x, ab)
        n = min(int(np.float32(nu) / np.float32(res) * (2 * np.pi)), n_rots)
        if n <= 0:
            continue
        ang = (np.arange(n) * 2 * np.pi / n).astype(np.float32).astype(np.float64)
        cand = c[None] + np.cos(ang)[:, None] * x[None] + np.sin(ang)[:, None] * y[None]
        d_tol = np.abs(np.linalg.norm(cand - centre[None], axis=1) - tol)
        g = cand - corner[None]
        d_face = np.minimum(np.abs(g), np.abs(g - hi[None])).min(1)
        near = np.linalg.norm(cand - centre[None], axis=1) <= tol * 1.001       # faces only matter for candidates inside the ball
        out[k] = min(d_tol.min(), d_face[near].min() if near.any() else np.inf)
    return out


def assert_masks_equal_up_to_rounding(got_mask, ref_mask, pc, tr, idxs, dims, corner, res, centre, tol, what, max_flips):
    """Bit-equality, or: every differing pair has a candidate within float32 rounding of a decision boundary, and there
    are at most `max_flips` of them.  Prints the exact count."""
    got_mask, ref_mask = np.asarray(got_mask, bool), np.asarray(ref_mask, bool)
    diff = np.nonzero(got_mask != ref_mask)[0]
    print(f"[{what}] mask bits differing: {len(diff)} of {len(ref_mask)}")
    if len(diff) == 0:
        return 0
    assert len(diff) <= max_flips, f"{what}: {len(diff)} mask bits differ (allowed: {max_flips})"
    scale = float(np.abs(pc).max() + np.abs(tr).max())
    margin = backvote_decision_margin(pc, tr, idxs, dims, corner, res, centre, tol, rows=diff)
    bound = 64 * np.finfo(np.float32).eps * scale                    # a few dozen float32 roundings of metre-scale values
    assert np.all(margin < bound), f"{what}: a differing pair is {margin.max():.3e} m from any decision boundary (> {bound:.3e})"
    return len(diff)
