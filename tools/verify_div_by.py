#!/usr/bin/env python
"""Exhaustive check of div_by (csrc/vote_common.cuh): q0 = a*y, r = fma(-b, q0, a), q = fma(r, y, q0) with y = RN(1/b)
against the IEEE quotient a / b, for EVERY float32 a in [2^lo, 2^hi) and the grid resolutions the reference ships
(config/category/*.yaml: 4e-3, 1e-2, 3e-2) plus the demo's 2e-2.  The FMAs are emulated in float64 (products of two
float32 are exact there; the one rounding to float32 is the FMA's).  Prints the number of differing quotients."""
import sys

import numpy as np


def check(res, lo=-30, hi=3):
    b = np.float32(res)
    y = np.float32(np.float32(1.0) / b)
    b64, y64 = np.float64(b), np.float64(y)
    bad = 0
    total = 0
    for e in range(lo, hi):
        bits = (np.arange(1 << 23, dtype=np.uint32) | np.uint32((e + 127) << 23))
        a = bits.view(np.float32)
        q0 = (a * y).astype(np.float32)
        r = (a.astype(np.float64) - b64 * q0.astype(np.float64)).astype(np.float32)      # fma(-b, q0, a)
        t = r.astype(np.float64) * y64 + q0.astype(np.float64)                            # exact product, one f64 add
        q = t.astype(np.float32)
        ref = (a / b).astype(np.float32)
        m = q.view(np.uint32) != ref.view(np.uint32)
        if m.any():
            # a float64 add followed by a float32 rounding can double-round: re-check those few in exact arithmetic
            from fractions import Fraction
            for i in np.nonzero(m)[0]:
                exact = Fraction(float(r[i])) * Fraction(float(y)) + Fraction(float(q0[i]))
                lo_f, hi_f = sorted((float(q[i]), float(ref[i])))
                # which float32 neighbour is nearest to the exact fma argument?
                cands = [np.float32(lo_f), np.float32(hi_f)]
                best = min(cands, key=lambda c: abs(Fraction(float(c)) - exact))
                if np.float32(best).view(np.uint32) != ref[i].view(np.uint32):
                    bad += 1
        total += a.size
    return bad, total


if __name__ == "__main__":
    for res in [float(v) for v in sys.argv[1:]] or [4e-3, 1e-2, 2e-2, 3e-2]:
        bad, total = check(res)
        print(f"res {res:g}: {bad} of {total} quotients differ from IEEE division")
