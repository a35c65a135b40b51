"""div_by (csrc/vote_common.cuh) replaces the reference's `/ res` (models/voting.py:35) by q0 = a*y, r = fma(-b, q0, a),
q = fma(r, y, q0) with y = RN(1/b).  tools/verify_div_by.py checks it against IEEE division for EVERY float32 in
[2^-30, 8) at the resolutions the reference ships (0 differences in 4 x 2.8e8 quotients); this is the quick version:
two full binades per resolution, through the same code."""
import importlib.util
import os

import pytest

_spec = importlib.util.spec_from_file_location(
    "verify_div_by", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "verify_div_by.py"))
verify = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(verify)


@pytest.mark.parametrize("res", [4e-3, 1e-2, 2e-2, 3e-2])
def test_div_by_equals_ieee_division_on_two_binades(res):
    bad, total = verify.check(res, lo=-4, hi=-2)          # candidate offsets of 6-25 cm: where the grids live
    assert total == 2 << 23 and bad == 0
