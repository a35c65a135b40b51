"""splat_fixed (csrc/vote_common.cuh) turns each trilinear corner weight into an integer number of 2^-14 units with one
FFMA onto 2^23: bits(fma(w, z * 2^14, 2^23)) - bits(2^23).  Restated here in numpy (the FMA's single rounding emulated
in float64, where the product of two float32 is exact): the integer is round-half-even(w * z * 2^14), and the eight
corner weights of a candidate sum to 2^14 within the eight roundings -- what makes the shared-memory grid an exact,
order-independent integer accumulator of models/voting.py:40-63."""
import numpy as np

F = np.float32
MAGIC = F(8388608.0)              # 2^23
MAGIC_BITS = np.uint32(0x4B000000)


def _fixed(w, z):
    s = (w.astype(np.float64) * z.astype(np.float64) + np.float64(MAGIC)).astype(F)      # fma(w, z, 2^23)
    return (s.view(np.uint32) - MAGIC_BITS).astype(np.int64)


def test_magic_number_rounding_is_round_half_even_of_the_scaled_weight():
    rng = np.random.default_rng(0)
    n = 2_000_000
    w = rng.random(n, dtype=F)                        # a product of two of (r, 1 - r)
    z = (rng.random(n, dtype=F) * F(16384.0)).astype(F)
    got = _fixed(w, z)
    exact = w.astype(np.float64) * z.astype(np.float64)
    assert np.array_equal(got, np.rint(exact).astype(np.int64))
    assert got.min() >= 0 and got.max() <= 16384


def test_eight_corner_weights_sum_to_one_within_the_roundings():
    rng = np.random.default_rng(1)
    n = 500_000
    g = (rng.random((n, 3), dtype=F) * F(50.0) + F(0.01)).astype(F)
    r = (g - np.floor(g)).astype(F)
    rx, ry, rz = r[:, 0], r[:, 1], r[:, 2]
    wx0, wy0 = (F(1) - rx).astype(F), (F(1) - ry).astype(F)
    z1, z0 = (rz * F(16384.0)).astype(F), ((F(1) - rz) * F(16384.0)).astype(F)
    total = np.zeros(n, np.int64)
    for wxy in ((wx0 * wy0).astype(F), (wx0 * ry).astype(F), (rx * wy0).astype(F), (rx * ry).astype(F)):
        for z in (z0, z1):
            total += _fixed(wxy, z)
    assert np.abs(total - 16384).max() <= 4           # eight half-unit roundings + the float32 products
    assert abs(float(total.mean()) - 16384.0) < 0.01  # unbiased


def test_denormal_product_carries_the_same_integer_in_its_bit_pattern():
    """The shipped formulation (csrc/vote_common.cuh, CPPF_SPLAT_DENORM): the x-y factor scaled by 2^-60 and the z factor by
    2^-75 (exact), so that their float32 product is a denormal whose BIT PATTERN is round-half-even(w * z * 2^14) -- the
    same integer as the magic-number FFMA, without the subtraction."""
    rng = np.random.default_rng(2)
    n = 2_000_000
    r = rng.random((n, 3), dtype=F)
    w = ((F(1) - r[:, 0]) * r[:, 1]).astype(F)                    # a product of two of (r, 1 - r)
    zf = np.where(rng.random(n) < 0.5, r[:, 2], (F(1) - r[:, 2]).astype(F)).astype(F)
    ws = (w * F(2.0 ** -60)).astype(F)                            # exact: normal range
    zs = (zf * F(2.0 ** -75)).astype(F)
    assert np.array_equal(ws.astype(np.float64), w.astype(np.float64) * 2.0 ** -60)
    assert np.array_equal(zs.astype(np.float64), zf.astype(np.float64) * 2.0 ** -75)
    prod = (ws * zs).astype(F)                                    # float32 multiply: one rounding, into the denormals
    got = prod.view(np.uint32).astype(np.int64)
    want = _fixed(w, (zf * F(16384.0)).astype(F))
    assert np.array_equal(got, want)
    # edge values: weight 0, weight 1, exact ties
    e_w = np.array([0.0, 1.0, 1.0, 0.5, 0.25, 3 * 2.0 ** -15, 2.0 ** -15], F)
    e_z = np.array([1.0, 1.0, 0.0, 2.0 ** -14, 2.0 ** -13, 1.0, 1.0], F)
    e = ((e_w * F(2.0 ** -60)).astype(F) * (e_z * F(2.0 ** -75)).astype(F)).astype(F).view(np.uint32).astype(np.int64)
    assert e.tolist() == [0, 16384, 0, 0, 0, 2, 0]               # half-even: 0.5 -> 0, 0.5 -> 0, 1.5 -> 2, 0.5 -> 0
