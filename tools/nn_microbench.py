#!/usr/bin/env python
"""som_assign / pairwise_min: cell-grid path against the brute-force kernels at the KITTI detector shape (16 clouds x 16384
points, 512 nodes / keypoints), on the bench's synthetic lidar clouds.  CUDA events, median of 20, L2 not flushed (both
variants read the same 3 MB)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import usip_oracle as orc
from usip_b200 import ops

dev = torch.device("cuda:0")
d = orc.synth_pair(8, 16384, 512, 4, kind="lidar", seed=1)
pc = torch.from_numpy(np.concatenate([d["src_pc"], d["dst_pc"]], 0)).to(dev).contiguous()
node = torch.from_numpy(np.concatenate([d["src_node"], d["dst_node"]], 0)).to(dev).contiguous()
kp = (node + 0.3 * torch.randn_like(node)).contiguous()


def timed(fn, n=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return round(float(np.median(ts)) * 1e3, 1)


out = {}
for m in ("brute", "grid"):
    out["som_assign_%s_us" % m] = timed(lambda: ops.som_assign(pc, node, method=m))
    out["pairwise_min_16x512x16384_%s_us" % m] = timed(lambda: ops.pairwise_min(kp, pc, method=m))
    out["pairwise_min_8x512x16384_%s_us" % m] = timed(lambda: ops.pairwise_min(kp[:8].contiguous(), pc[:8].contiguous(), method=m))
ig, cg = ops.som_assign(pc, node, method="grid"); ib, cb = ops.som_assign(pc, node, method="brute")
out["som_assign_equal"] = bool(torch.equal(ig, ib) and torch.equal(cg, cb))
dg, ag = ops.pairwise_min(kp, pc, method="grid"); db, ab = ops.pairwise_min(kp, pc, method="brute")
out["pairwise_min_equal"] = bool(torch.equal(ag, ab) and torch.equal(dg, db))
print(json.dumps(out, indent=1))
