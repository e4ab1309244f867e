"""Drop-in for the reference's `index_max` extension module (models/index_max_ext/index_max.cpp:154-159).

    forward_cuda_shared_mem(data, index, K) / forward_cuda(data, index, K) -> int32 (B,C,K)

Both names run the same sm_100a kernel (usip_index_max_f32): the reference's two variants only differ in
where the running max lives, and its shared-memory variant silently returns zeros once B*K > 12288
(index_max_cuda.cu:92-96, no cudaFuncSetAttribute) -- that limit does not exist here.
The CPU entry points are deliberately not provided: this build has no CPU path."""
import os as _os
import sys as _sys

# importable both as `usip_b200.index_max` and -- with <repo>/usip_b200 first on sys.path, the reference's own import style
# (models/networks.py:9-18) -- as the bare top-level name: make the `usip_b200` package itself resolvable either way
_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _ROOT not in _sys.path:
    _sys.path.append(_ROOT)
from usip_b200 import ops as _ops  # noqa: E402


def forward_cuda_shared_mem(data, index, K):
    return _ops.index_max(data, index, K)


def forward_cuda(data, index, K):
    return _ops.index_max(data, index, K)


def forward_cpu(data, index, K):
    raise NotImplementedError("usip_b200.index_max is GPU-only (no CPU fallback); the CPU restatement of "
                              "index_max.cpp:73-112 lives in oracle/ as test infrastructure")


def forward_multi_thread_cpu(data, index, K, thread_num):
    raise NotImplementedError("usip_b200.index_max is GPU-only (no CPU fallback)")
