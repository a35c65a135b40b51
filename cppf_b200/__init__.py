"""cppf_b200 -- B200-native (sm_100a) implementation of CPPF's per-object hot path:
point-pair features -> pair MLP -> centre / orientation voting.

The compute lives in hand-written CUDA behind a C ABI (``include/cppf_b200.h``,
``cppf_b200/csrc``); this package is the thin Python host side that mirrors the
reference's ``models/model.py`` and ``models/voting.py`` surface.  There is no CPU
or PyTorch fallback: using any operator without the built ``libcppf_b200.so`` raises.
"""
__version__ = "0.1.0"
