"""Drop-in for the reference's `ball_query` extension module (models/ball_query_ext/ball_query.cpp:45-48).

    forward_cuda_shared_mem(node_to_point_dist, radius, K) -> int32 (B,M,K)

plus `forward_fused(xyz, feat, centers, radius, K)`, the B200 path that never builds the (B,M,N) distance
matrix (replaces models/networks.py:355-373)."""
import os as _os
import sys as _sys

# importable both as `usip_b200.ball_query` and -- with <repo>/usip_b200 first on sys.path, the reference's own import style
# (models/networks.py:9-18) -- as the bare top-level name: make the `usip_b200` package itself resolvable either way
_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _ROOT not in _sys.path:
    _sys.path.append(_ROOT)
from usip_b200 import ops as _ops  # noqa: E402


def forward_cuda_shared_mem(node_to_point_dist, radius, K):
    return _ops.ball_query_dist(node_to_point_dist, radius, K)


def forward_cuda(node_to_point_dist, radius, K):
    # ball_query.cpp:23-31 prints "Not implemented yet." and falls off a non-void function (UB).
    raise NotImplementedError("ball_query.forward_cuda is not implemented in the reference either "
                              "(ball_query.cpp:23-31); use forward_cuda_shared_mem")


def forward_fused(xyz, feat, centers, radius, K, want_group=True):
    return _ops.ball_group(xyz, feat, centers, radius, K, want_group)
