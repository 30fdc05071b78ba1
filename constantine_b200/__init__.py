"""constantine_b200 -- B200 (sm_100a) multi-scalar-multiplication engine behind Constantine's C ABI.

Only the MSM hot path lives here: csrc/ (CUDA kernels + the extern "C" boundary, built into lib/libctt_b200_msm.so),
curves.py (curve metadata) and msm.py (host-side mirror of the reference's MSM interface over that C ABI).
"""
from .curves import CURVES, CurveParams  # noqa: F401
from . import msm  # noqa: F401
