"""Mirror of the live part of models/operations.py: the kNN gather (`knn_gather_by_indexing`, `knn_gather_wrapper`,
operations.py:243-287) and the two shared-memory constants callers read.  The ~250 lines of commented-out Numba kernels
of the reference are the ancestors of index_max / ball_query and are served by usip_b200/index_max.py, ball_query.py.

Inside the fused detector plan the gather never materialises (B,C,N,K) (the consuming layer gathers its operand rows);
these functions are the stand-alone entry points with the reference's signature."""
import torch

from usip_b200 import ops

# generalized batch size / SOM size limits of the reference's shared-memory kernels (operations.py:16-18); kept for
# attribute compatibility -- the B200 kernels have no such cap.
CUDA_SHARED_MEM_DIM_X = 24
CUDA_SHARED_MEM_DIM_Y = 512


def knn_gather_by_indexing(som_node, som_node_knn_I):
    """som_node (B,C,N), som_node_knn_I (B,N,K) -> (B,C,N,K): out[b,c,n,k] = som_node[b,c,I[b,n,k]]."""
    if torch.is_grad_enabled() and som_node.requires_grad:
        raise NotImplementedError("stand-alone knn_gather has no autograd; it is fused inside the network-level plan")
    return ops.knn_gather(som_node.contiguous().float(), som_node_knn_I.contiguous().to(torch.int32))


def knn_gather_wrapper(som_node, som_node_knn_I):
    assert som_node.size()[1] == 3
    return knn_gather_by_indexing(som_node, som_node_knn_I)
