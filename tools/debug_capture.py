#!/usr/bin/env python
"""Find which stage of the train step breaks CUDA-graph capture (debug aid)."""
import os, sys, traceback
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import usip_oracle as orc
from tests.util_gpu import load_params, make_opt
from usip_b200 import _lib
from usip_b200.models.keypoint_detector import ModelDetector

B, N, M, S, Kn = 2, 2048, 64, 4, 16
def mk():
    opt = make_opt(batch_size=B, input_pc_num=N, node_num=M, surface_normal_len=S, node_knn_k_1=Kn)
    md = ModelDetector(opt)
    load_params(md.detector, orc.init_detector_params(S=S, seed=0, randomize_bn=True))
    d = orc.synth_pair(B, N, M, S, kind="lidar", seed=1)
    md.set_input(*[torch.from_numpy(d[k]) for k in ("src_pc", "src_sn", "src_node", "dst_pc", "dst_sn", "dst_node", "R", "scale", "shift")])
    md.detector.train()
    return md

def stage_fwd(md): md._run_siamese(is_train=True, epoch=0)
def stage_direct(md): md._train_step_direct(0)
def stage_all(md): md._optimize_eager(0)

for mode in ("global",):
    for name, fn in (("fwd(keep)", stage_fwd), ("direct fwd+loss+bwd", stage_direct), ("+adam", stage_all)):
        md = mk()
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2): fn(md)
        torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
        _lib.WEIGHT_GEN[0] += 1
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g, capture_error_mode=mode):
                fn(md)
            g.replay(); torch.cuda.synchronize()
            print("[%s] %-10s capture+replay OK" % (mode, name), flush=True)
        except Exception as e:
            print("[%s] %-10s FAILED: %s" % (mode, name, str(e).splitlines()[0]), flush=True)
            traceback.print_exc()
            sys.exit(0 if mode == "thread_local" else 1) if False else None
            break
