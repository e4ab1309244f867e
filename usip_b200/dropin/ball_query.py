"""Operator-level drop-in for the reference's `ball_query` extension module (see dropin/index_max.py)."""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _ROOT not in _sys.path:
    _sys.path.append(_ROOT)
from usip_b200.ball_query import forward_cuda, forward_cuda_shared_mem, forward_fused  # noqa: E402,F401
