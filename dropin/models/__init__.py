"""`models` package with the reference's module names, backed by cppf_b200.

Put ``<repo>/dropin`` (and ``<repo>``) on ``sys.path`` ahead of the reference tree and
``from models.model import PPFEncoder, PointEncoder`` /
``from models.voting import rot_voting_kernel, backvote_kernel, ppf_kernel``
(nocs/inference.py:3,17) resolve to the sm_100a implementation.
"""
