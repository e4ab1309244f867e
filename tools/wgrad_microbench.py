#!/usr/bin/env python
"""usip_wgrad at the train step's shapes: time (CUDA events, median of 10) for precision 1 (3xTF32) and 4 (single TF32),
max error against an fp64 matmul, and the shared-memory / tensor floors of the 3xTF32 kernel."""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from usip_b200 import _lib

dev = torch.device("cuda:0")
lib = _lib.load()
if len(sys.argv) > 1:                                   # A/B: time another build of the library (same ABI)
    lib = ctypes.CDLL(sys.argv[1])
    lib.usip_wgrad.restype = ctypes.c_int
    lib.usip_wgrad.argtypes = _lib.SIGNATURES["usip_wgrad"][1]
P_ = lambda t: ctypes.c_void_p(t.data_ptr())
out = {}
for (P, Cin, Cout) in [(131072, 512, 512), (131072, 256, 512), (131072, 256, 256), (131072, 128, 256), (262144, 128, 128),
                       (262144, 64, 128), (262144, 64, 64), (65536, 256, 256), (8192, 512, 512)]:
    gy = torch.randn(P, Cout, device=dev); x = torch.randn(P, Cin, device=dev)
    sc = torch.rand(Cin, device=dev) + 0.5; sh = torch.randn(Cin, device=dev)
    ref = gy.double().t() @ torch.relu(x.double() * sc.double() + sh.double())
    res = {}
    for prec in (1, 4):
        gW = torch.zeros(Cout, Cin, device=dev)
        def run():
            _lib.check(lib.usip_wgrad(P_(gy), Cout, P_(x), Cin, P_(sc), P_(sh), 1, P_(gW), Cin, P, Cout, Cin, prec,
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "usip_wgrad")
        run(); torch.cuda.synchronize()
        err = float(((gW.double() - ref).abs().max() / ref.abs().max()).item())
        for _ in range(3): run()
        ts = []
        for _ in range(10):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res["p%d_us" % prec] = round(float(np.median(ts)) * 1e3, 1)
        res["p%d_err" % prec] = float("%.2e" % err)
    flops = 2.0 * P * Cin * Cout
    res["tflops_alg_p1"] = round(flops / res["p1_us"] / 1e6, 1)
    res["mma_floor_us_at_1965MHz"] = round(3 * flops / (148 * 2048 * 2 * 1.965e9) * 1e6, 1)
    out["%dx%d->%d" % (P, Cin, Cout)] = res
print(json.dumps(out, indent=1))
