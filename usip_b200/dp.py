"""Data-parallel plumbing: ONE flat fp32 gradient buffer, ONE all-reduce per step.

The reference uses nn.DataParallel (models/keypoint_detector.py:35-37): single process, parameters broadcast every
forward, gradients reduce-added to device 0.  Here the unit is one process per GPU (torch.distributed); the only
exchange on the path is the gradient sum, so all parameter `.grad`s are made views into one contiguous buffer
(1,198,724 floats = 4.79 MB for RPN_Detector) that is all-reduced once (NCCL over NVLink/NVSwitch on the GPU box,
gloo in the CPU tests).  BatchNorm statistics stay per rank, which is exactly DataParallel's per-replica semantics.
"""
import gc
import os
import threading

import torch
import torch.distributed as dist


class FlatGradAllReduce:
    """`params` is either an iterable of parameters (a flat gradient buffer is built here) or a usip_b200.optim.FlatAdam,
    whose flat parameter / gradient / moment buffers are used as they are (one broadcast each, one all-reduce per step)."""

    def __init__(self, params, process_group=None, broadcast_from=0, buffers=()):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = process_group if process_group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.avg_op = dist.get_backend(self.group) == "nccl"          # ncclAvg: sum and 1/world in the one collective
        if hasattr(params, "flat_g") and hasattr(params, "flat_p"):  # FlatAdam
            opt = params
            self.params = list(opt._params)
            self.flat = opt.flat_g
            for t in [opt.flat_p, opt.exp_avg, opt.exp_avg_sq, opt.step_dev] + list(buffers):
                dist.broadcast(t.data, src=broadcast_from, group=self.group)
            return
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        n = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(n, dtype=p0.dtype, device=p0.device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)     # autograd accumulates in place into the view
            off += p.numel()
        # identical parameters / buffers on every rank
        for t in list(self.params) + list(buffers):
            dist.broadcast(t.data, src=broadcast_from, group=self.group)

    def zero(self):
        self.flat.zero_()

    def allreduce_mean(self):
        if self.avg_op:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / self.world)

    def check_views(self):
        """True iff every .grad still aliases the flat buffer (guards against zero_grad(set_to_none=True))."""
        base = self.flat.data_ptr()
        end = base + self.flat.numel() * self.flat.element_size()
        return all(p.grad is not None and base <= p.grad.data_ptr() < end for p in self.params)


def shutdown(*models, hard_exit_after=None):
    """Orderly end of a data-parallel job: the models' CUDA graphs (which hold the captured NCCL all-reduce) are destroyed
    first, then the device is drained, then the process group.  Destroying the communicator under a live graph deadlocks
    (seen on 2xB200: both ranks hung in destroy_process_group after a finished run).  `hard_exit_after` (seconds) arms a
    watchdog that ends the process with status 0 if the teardown itself does not return -- for benchmark / test drivers
    whose results are already written."""
    for m in models:
        if hasattr(m, "release_cuda_graphs"):
            m.release_cuda_graphs()
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if not dist.is_initialized():
        return
    timer = None
    if hard_exit_after:
        timer = threading.Timer(float(hard_exit_after), lambda: os._exit(0))
        timer.daemon = True
        timer.start()
    try:
        dist.barrier()
        dist.destroy_process_group()
    finally:
        if timer is not None:
            timer.cancel()
