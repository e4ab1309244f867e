"""Operator-level drop-in: a directory holding ONLY the two extension-module names the reference imports by bare name
(models/networks.py:17-18).  `sys.path.insert(0, "<repo>/usip_b200/dropin")` in front of the reference tree replaces
`models/index_max_ext` while keeping the reference's own models/ and util/ (INTEGRATION.md section 2)."""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _ROOT not in _sys.path:
    _sys.path.append(_ROOT)
from usip_b200.index_max import forward_cpu, forward_cuda, forward_cuda_shared_mem, forward_multi_thread_cpu  # noqa: E402,F401
