"""-m gpu, needs >= 2 GPUs (skipped otherwise): ModelDetector.enable_data_parallel() on NCCL, one process per GPU --
the replacement of nn.DataParallel (models/keypoint_detector.py:35-37), BASELINE.json configs[4] in miniature.

Two ranks start from DIFFERENT parameters (the broadcast must fix that), each runs optimize() on its own pairs:
  * parameters and BN-independent state are bit-identical on both ranks after every step;
  * the all-reduced gradient equals the mean of the two per-rank gradients computed by single-process runs
    (per-rank BatchNorm statistics = DataParallel's per-replica semantics);
  * the parameters after the step equal a single-process Adam step on that mean gradient."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
CFG = dict(B=2, N=4096, M=128, S=4, Kn=16)
KEYS = ("src_pc", "src_sn", "src_node", "dst_pc", "dst_sn", "dst_node", "R", "scale", "shift")


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _make(rank_seed, dev):
    sys.path.insert(0, ROOT)
    from oracle import usip_oracle as orc
    from tests.util_gpu import load_params, make_opt
    from usip_b200.models.keypoint_detector import ModelDetector
    opt = make_opt(batch_size=CFG["B"], input_pc_num=CFG["N"], node_num=CFG["M"], surface_normal_len=CFG["S"],
                   node_knn_k_1=CFG["Kn"], device=dev, gpu_ids=[dev.index])
    md = ModelDetector(opt)
    P = orc.init_detector_params(S=CFG["S"], seed=rank_seed, randomize_bn=True)
    P["mlp3.conv.weight"] = (P["mlp3.conv.weight"] * 1000).astype(np.float32)
    load_params(md.detector, P)
    return md, orc


def _batch(orc, rank, step):
    d = orc.synth_pair(CFG["B"], CFG["N"], CFG["M"], CFG["S"], kind="lidar", seed=500 + 10 * rank + step)
    return [torch.from_numpy(d[k]) for k in KEYS]


def worker(out_dir):
    import torch.distributed as dist
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    md, orc = _make(rank_seed=7 + rank, dev=dev)            # rank 1 starts from other weights
    md.enable_data_parallel()
    rec = {}
    for step in range(2):
        md.set_input(*_batch(orc, rank, step))
        md.optimize(epoch=0)
        torch.cuda.synchronize()
        rec["grad%d" % step] = [p.grad.detach().cpu().clone() for p in md.detector.parameters()]
        rec["param%d" % step] = [p.detach().cpu().clone() for p in md.detector.parameters()]
        rec["loss%d" % step] = float(md.loss)
    torch.save(rec, os.path.join(out_dir, "rank%d.pt" % rank))
    from usip_b200.dp import shutdown
    shutdown(md, hard_exit_after=20)          # graphs (captured all-reduce) before the communicator


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_enable_data_parallel_two_ranks_nccl(tmp_path):
    port = _free_port()
    env = {**os.environ, "PYTHONPATH": ROOT}
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__), "--worker", str(tmp_path)],
                       env=env, capture_output=True, text=True, timeout=420)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    a = torch.load(os.path.join(tmp_path, "rank0.pt")); b = torch.load(os.path.join(tmp_path, "rank1.pt"))
    for step in range(2):
        for x, y in zip(a["param%d" % step], b["param%d" % step]):
            assert torch.equal(x, y)                         # replicas stay bit-identical
        for x, y in zip(a["grad%d" % step], b["grad%d" % step]):
            assert torch.equal(x, y)                         # both hold the reduced gradient
        assert a["loss%d" % step] != b["loss%d" % step]      # ... of different batches

    # single-process reference: same start (rank 0's weights), per-rank gradients one after the other, mean, Adam
    dev = torch.device("cuda", 0)
    md, orc = _make(rank_seed=7, dev=dev)
    for step in range(2):
        start = [p.detach().clone() for p in md.detector.parameters()]
        bufs = [bf.detach().clone() for bf in md.detector.buffers()]
        grads = []
        for rank in range(2):
            with torch.no_grad():
                for p, s0 in zip(md.detector.parameters(), start):
                    p.copy_(s0)
                for bf, b0 in zip(md.detector.buffers(), bufs):
                    bf.copy_(b0)
            md.set_input(*_batch(orc, rank, step))
            md.detector.train()
            md._run_siamese(is_train=True, epoch=0)
            md.detector.zero_grad()
            md._losses()
            md.loss.backward()
            grads.append([p.grad.detach().clone() for p in md.detector.parameters()])
        mean = [(g0 + g1) * 0.5 for g0, g1 in zip(*grads)]
        gmax = max(float(m.abs().max()) for m in mean)
        for k, (m, got) in enumerate(zip(mean, a["grad%d" % step])):
            scale = max(float(m.abs().max()), 1e-12)
            # conv biases in front of a BatchNorm have an analytically zero gradient: what is left there is summation
            # noise (1e-7 of the largest gradient), so the bound has a floor relative to the whole gradient
            assert float((m.cpu() - got).abs().max()) <= 2e-5 * scale + 1e-6 * gmax, (step, k, scale, gmax)
        # continue the single-process trajectory from the data-parallel parameters (Adam state is per process)
        with torch.no_grad():
            for p, v in zip(md.detector.parameters(), a["param%d" % step]):
                p.copy_(v.to(dev))
        if step == 0:
            # first Adam step from identical start: lr * g / (|g| + eps)
            lr = md.opt.lr
            for s0, m, v in zip(start, mean, a["param0"]):
                want = s0 - lr * m / (m.abs() + 1e-8)
                if float(m.abs().max()) < 1e-5 * gmax:      # analytic-zero gradient (noise only): sign(g) is arbitrary
                    continue
                solid = m.abs() > 1e-3 * m.abs().max()
                assert float((want.cpu() - v).abs()[solid.cpu()].max()) <= 0.02 * lr + 1e-7


if __name__ == "__main__" and len(sys.argv) >= 3 and sys.argv[1] == "--worker":
    sys.path.insert(0, ROOT)
    worker(sys.argv[2])
