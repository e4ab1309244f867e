#!/usr/bin/env python
"""Node generation: GPU farthest point sampling vs the numpy FarthestSampler restatement on the host,
KITTI shape: 16 clouds x 5461 candidate points (N/3) x 512 nodes."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from usip_b200 import ops

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
B, Ns, k = 16, 5461, 512
pts = (rng.uniform(-40, 40, (B, Ns, 3)) * np.array([1, 0.05, 1])).astype(np.float32)
start = rng.integers(0, Ns, B).astype(np.int32)
P = torch.from_numpy(pts).to(dev); S = torch.from_numpy(start).to(dev)
for _ in range(3): ops.fps(P, S, k)
ts = []
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); ops.fps(P, S, k); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
def numpy_fps(p, st, kk):                       # the reference's algorithm (float64 distances, np.argmax)
    far = np.zeros((kk, 3)); far[0] = p[st]
    d = ((far[0] - p) ** 2).sum(axis=1)
    for i in range(1, kk):
        far[i] = p[np.argmax(d)]
        d = np.minimum(d, ((far[i] - p) ** 2).sum(axis=1))
    return far
t0 = time.perf_counter(); numpy_fps(pts[0], int(start[0]), k); t_np = time.perf_counter() - t0
print(json.dumps({"workload": "FPS 16 clouds x 5461 points -> 512 nodes", "gpu_ms_16_clouds": float(np.median(ts)),
                  "numpy_ms_per_cloud_1_core": t_np * 1e3, "speedup_vs_serial_numpy_16_clouds": t_np * 1e3 * B / float(np.median(ts))}))
