import os as _os
import sys as _sys

# importable both as `usip_b200.models` and -- with <repo>/usip_b200 first on sys.path, the reference's own import style
# (models/networks.py:9-18) -- as the bare top-level name: make the `usip_b200` package itself resolvable either way
_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _ROOT not in _sys.path:
    _sys.path.append(_ROOT)
