"""Drop-in for the reference's models/voting.py."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from cppf_b200.voting import backvote_kernel, findpeak_kernel, ppf_kernel, rot_voting_kernel  # noqa: E402,F401
