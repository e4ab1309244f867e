"""World-size-2 gloo test (CPU) of the data-parallel plumbing: flat gradient buffer, single all-reduce, identical
parameters after an Adam step on both ranks, equal to a single-process run on the concatenated batch when the model
has no batch statistics."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from usip_b200.dp import FlatGradAllReduce
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                       # different init per rank: broadcast must fix it
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    dp = FlatGradAllReduce(net.parameters())
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    g = torch.Generator().manual_seed(7)
    X = torch.randn(8, 6, generator=g); Y = torch.randn(8, 3, generator=g)
    xs, ys = X[rank * 4:(rank + 1) * 4], Y[rank * 4:(rank + 1) * 4]          # each rank keeps its own pairs
    for _ in range(3):
        dp.zero()
        loss = ((net(xs) - ys) ** 2).mean()
        loss.backward()
        assert dp.check_views()
        dp.allreduce_mean()
        opt.step()
    torch.save([p.detach().clone() for p in net.parameters()], os.path.join(out_dir, "rank%d.pt" % rank))

    class _Holder:                                       # stands in for a model that holds captured step graphs
        released = False

        def release_cuda_graphs(self):
            self.released = True
    from usip_b200.dp import shutdown
    h = _Holder()
    shutdown(h, hard_exit_after=60)                      # graphs first, then barrier + destroy_process_group
    assert h.released and not dist.is_initialized()


def test_flat_grad_allreduce_two_ranks(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(os.path.join(tmp_path, "rank0.pt")); b = torch.load(os.path.join(tmp_path, "rank1.pt"))
    for x, y in zip(a, b):
        assert torch.equal(x, y)                                                # replicas stay bit-identical
    # single-process reference on the full batch (mean of per-rank means == full mean for equal shards)
    torch.manual_seed(100)
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    g = torch.Generator().manual_seed(7)
    X = torch.randn(8, 6, generator=g); Y = torch.randn(8, 3, generator=g)
    for _ in range(3):
        opt.zero_grad()
        ((net(X) - Y) ** 2).mean().backward()
        opt.step()
    for x, y in zip(a, net.parameters()):
        assert torch.allclose(x, y.detach(), rtol=1e-5, atol=1e-6)
