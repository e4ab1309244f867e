"""Mirror of the one hot-path function of util/som.py: `query_topk` (util/som.py:17-54).

    mask, mask_row_max, min_idx = query_topk(node, x, M, k)

Only k == 1 is implemented (every shipped config; networks.py:85).  The nearest-node search runs in the
sm_100a kernel usip_som_assign_f32 without ever building the (B,C,N,M) difference tensor; the dense one-hot
`mask` (B,N,M) int32 the reference returns is materialised here ONLY for API compatibility -- the fused
network plan (usip_b200/engine.py) consumes min_idx / counts directly and never calls this function.
The SOM / BatchSOM fitting classes (util/som.py:57-417) are not on the hot path and are not provided."""
import torch

from usip_b200 import ops


def query_topk(node, x, M, k):
    if k != 1:
        raise NotImplementedError("query_topk: only k=1 is implemented on the B200 path")
    node = node.to(x.device)
    min_idx32, count = ops.som_assign(x.detach().contiguous().float(), node.detach().contiguous().float())
    min_idx = min_idx32.long()                                              # (B, kN)
    B, N = min_idx.shape
    mask = torch.zeros((B, N, M), dtype=torch.int32, device=x.device)
    mask.scatter_(2, min_idx.unsqueeze(2), 1)                               # one-hot, API compatibility only
    mask_row_max = (count > 0).to(torch.int32)                              # (B, M)
    return mask, mask_row_max, min_idx
