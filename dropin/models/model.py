"""Drop-in for the reference's models/model.py."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from cppf_b200.model import PPFEncoder, PointEncoder, ResLayer  # noqa: E402,F401
