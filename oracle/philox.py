"""TEST INFRASTRUCTURE (oracle) -- numpy restatement of the product's counter-based RNG.

Philox4x32-10 (Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3",
SC'11), the published algorithm; this is not reference code (the reference is unseeded,
SURVEY.md section 4) but the definition the product kernels' `seed` mode follows
(cppf_b200/csrc/common.cuh: philox4x32_10, u01).
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c = [np.asarray(v, dtype=np.uint32).copy() for v in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = np.uint32(k0), np.uint32(k1)
    for _ in range(10):
        p0 = _M0 * c[0].astype(np.uint64)
        p1 = _M1 * c[2].astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        with np.errstate(over="ignore"):
            k0, k1 = np.uint32(k0 + _W0), np.uint32(k1 + _W1)
    return c


def u01(w):
    return ((w >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def pair_uniforms(seed, n_pairs):
    """[n_pairs,4] uniforms of the fused encode+sample kernel: counter (p_lo, p_hi, 0, 0), key = seed."""
    p = np.arange(n_pairs, dtype=np.uint64)
    w = philox4x32_10((p & np.uint64(0xFFFFFFFF)).astype(np.uint32), (p >> np.uint64(32)).astype(np.uint32), 0, 0,
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack([u01(x) for x in w], -1)


def pair_uniforms_at(seed, pair_index):
    """pair_uniforms for an arbitrary set of pair indices (int64 array) -- the counter IS the pair index."""
    p = np.asarray(pair_index).astype(np.uint64)
    w = philox4x32_10((p & np.uint64(0xFFFFFFFF)).astype(np.uint32), (p >> np.uint64(32)).astype(np.uint32), 0, 0,
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack([u01(x) for x in w], -1)


def row_uniforms(seed, n_rows, stream_id):
    """uniforms of cppf_sample_bins mode 2: counter (row_lo, row_hi, stream_id, 0)."""
    p = np.arange(n_rows, dtype=np.uint64)
    w = philox4x32_10((p & np.uint64(0xFFFFFFFF)).astype(np.uint32), (p >> np.uint64(32)).astype(np.uint32), stream_id, 0,
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return u01(w[0])
