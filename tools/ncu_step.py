#!/usr/bin/env python
"""Target of ONE `ncu --set full` pass over every kernel of the hot path at the BASELINE shapes (tools/gpu_ncu_step.sh):
inside the cudaProfilerStart/Stop range run, eagerly, one KITTI detector fwd+loss, one optimize() (backward + Adam), the
Oxford fused ball-query+group, the stand-alone index_max / ball_query operators and one descriptor forward.
Everything before the range (two warm-up steps) is not profiled (`ncu --profile-from-start off`)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import usip_oracle as orc          # synthetic inputs / parameter init only
from tests.util_gpu import load_params, make_opt
from usip_b200 import ball_query, index_max, ops
from usip_b200.models import networks
from usip_b200.models.keypoint_detector import ModelDetector

what = set(sys.argv[1:]) or {"fwd", "train", "desc", "ops"}
dev = torch.device("cuda:0")
B, N, M, S, Kn = 8, 16384, 512, 4, 16
opt = make_opt(batch_size=B, input_pc_num=N, node_num=M, surface_normal_len=S, node_knn_k_1=Kn)
md = ModelDetector(opt)
load_params(md.detector, orc.init_detector_params(S=S, seed=0))
d = orc.synth_pair(B, N, M, S, kind="lidar", seed=1236)
md.set_input(*[torch.from_numpy(d[k]) for k in ("src_pc", "src_sn", "src_node", "dst_pc", "dst_sn", "dst_node", "R", "scale", "shift")])

Bd, Md, K = 16, 1024, 64
torch.manual_seed(1237)
pc = torch.stack([torch.empty(Bd, N, device=dev).uniform_(-40, 40), torch.empty(Bd, N, device=dev).uniform_(-2, 2),
                  torch.empty(Bd, N, device=dev).uniform_(-40, 40)], 1).contiguous()
sn = torch.randn(Bd, S, N, device=dev)
sel = torch.randint(0, N, (Bd, Md), device=dev)
kp = (torch.gather(pc, 2, sel.unsqueeze(1).expand(Bd, 3, Md)) + 0.1 * torch.randn(Bd, 3, Md, device=dev)).contiguous()
dopt = make_opt(batch_size=Bd // 2, input_pc_num=N, node_num=Md, surface_normal_len=S, ball_radius=1.0, ball_nsamples=K)
dnet = networks.DescriptorLiteOld(dopt).to(dev)
data = torch.randn(16, 128, N, device=dev)
index = torch.randint(0, 512, (16, N), device=dev, dtype=torch.int32)
dist = torch.norm(kp[:4].unsqueeze(3) - pc[:4].unsqueeze(2), dim=1).contiguous()


def run():
    if "fwd" in what:
        md.forward_loss(epoch=0, train_bn=True, graph=False)
    if "train" in what:
        md.optimize(epoch=0)
    if "desc" in what:
        ops.ball_group(pc, sn, kp, 1.0, K, want_group=True)
        dnet.train()
        with torch.no_grad():
            dnet(pc, sn, kp, True, None)
    if "ballonly" in what:
        ops.ball_group(pc, sn, kp, 1.0, K, want_group=True)
        index_max.forward_cuda_shared_mem(data, index, 512)
    if "ops" in what:
        index_max.forward_cuda_shared_mem(data, index, 512)
        ball_query.forward_cuda_shared_mem(dist, 1.0, K)


for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
if "ballonly" in what:
    for _ in range(3):
        run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ncu_step done:", sorted(what))
