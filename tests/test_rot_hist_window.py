"""The fused orientation histogram (csrc/vote_private.cu, rot_hist_kernel) does not multiply every candidate with every
sphere bin like nocs/inference.py:276-284 ([720 000, 3] x [3, 480], count dot > cos 1.5 deg): a candidate c and a bin s with
c . s > thr are closer than sqrt(2 - 2 thr), hence so are their y coordinates, and the Fibonacci sphere of
utils/util.py:102-118 has y strictly decreasing in the bin index -- so only the bins of a y window can count.  This CPU test
restates the window (the half-width the launcher computes, the two binary searches) and checks on many random candidates
that the windowed count equals the full scan, bin for bin, with the same float32 dot-product expression."""
import numpy as np

from cppf_b200.pipeline import fibonacci_sphere

F = np.float32


def test_y_window_scan_counts_exactly_what_the_full_scan_counts():
    sphere = fibonacci_sphere(480).astype(F)
    assert np.all(np.diff(sphere[:, 1]) < 0)                                   # y strictly decreasing: the kernel's `mono` test
    thr = F(np.cos(1.5 / 180 * np.pi))
    ywin = F(np.sqrt(max(0.0, 2.00002 - 2.0 * float(thr))) + 1e-4)              # rot_hist_launch
    rng = np.random.default_rng(0)
    n = 200_000
    c = rng.standard_normal((n, 3))
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    c[: n // 4] = sphere[rng.integers(0, 480, n // 4)] + rng.normal(0, 0.01, (n // 4, 3))      # near bin centres: real hits
    c = (c / (np.linalg.norm(c, axis=1, keepdims=True) + 1e-7)).astype(F)      # models/voting.py:144 normalisation
    c[:100] = 0                                                                # degenerate pairs leave the zero vector
    dots = ((c[:, None, 0] * sphere[None, :, 0]).astype(F) + (c[:, None, 1] * sphere[None, :, 1]).astype(F)
            + (c[:, None, 2] * sphere[None, :, 2]).astype(F))                  # rounding order is immaterial to the claim
    full = dots > thr
    ys = sphere[:, 1]
    y_hi, y_lo = (c[:, 1] + ywin).astype(F), (c[:, 1] - ywin).astype(F)
    lo = (ys[None, :] > y_hi[:, None]).sum(1)                                  # first s with y_s <= y_hi
    hi = (ys[None, :] >= y_lo[:, None]).sum(1)                                 # first s with y_s <  y_lo
    idx = np.arange(480)[None, :]
    inside = (idx >= lo[:, None]) & (idx < hi[:, None])
    assert not (full & ~inside).any()                                          # nothing that counts lies outside the window
    np.testing.assert_array_equal((full & inside).sum(0), full.sum(0))         # identical histogram
    assert full.sum() > 10_000                                                 # the case is not vacuous
    assert inside.sum(1).mean() < 40                                           # and the window is narrow (480 bins in the full scan)
