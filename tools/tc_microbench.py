#!/usr/bin/env python
"""Where does the time of the wide tcgen05 layer go?  Times the 512->512 / 256->256 layers on 131072 rows with parts of
the kernel disabled through usip_layer_desc.debug_flags (results are wrong in those runs; timing only)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from usip_b200 import ops

dev = torch.device("cuda:0")
out = {}
for (P, Cin, Cout, group, write_y) in [(131072, 512, 512, 16, False), (131072, 256, 256, 0, True), (131072, 256, 512, 0, True),
                                        (262144, 128, 128, 0, True), (262144, 64, 64, 0, True)]:
    X = torch.randn(P, Cin, device=dev); W = torch.randn(Cout, Cin, device=dev) / Cin ** 0.5
    b = torch.randn(Cout, device=dev); sc = torch.rand(Cin, device=dev) + 0.5; sh = torch.randn(Cin, device=dev)
    Y = torch.empty(P, Cout, device=dev) if write_y else None
    part = torch.empty(ops.stat_slots(P, Cout, 1, group, bool(group)), 2, Cout, device=dev)
    ws = torch.empty(2 * Cin * Cout, device=dev)
    kw = {}
    if group:
        Q = P // group
        kw = dict(gmax=torch.empty(Q, Cout, device=dev), gmin=torch.empty(Q, Cout, device=dev), group=group)
    res = {}
    for name, flags in [("full", 0), ("no_epilogue", 1), ("no_xload", 2), ("no_wtma", 4), ("no_Ystore", 32), ("no_stats_loop", 64),
                        ("no_epi+no_xload+no_wtma", 7), ("tf32+2xbf16 cross terms", 16), ("with L2 prefetch of X", 128)]:
        def run(packed):
            ops.layer_fwd(X, W, b, P, Cin, Cout, in_scale=sc, in_shift=sh, in_relu=True, Y=Y, stat_partial=part,
                          precision=1, tc_ws=ws, tc_packed=packed, debug_flags=flags, **kw)
        run(False)
        for _ in range(3): run(True)
        ts = []
        for _ in range(10):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); run(True); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res[name] = round(float(np.median(ts)) * 1e3, 1)
    # per-role time attribution of the full kernel (clock64 accumulators, averaged over CTAs / warps of a role)
    clk = torch.zeros(148 * 17 * 8, dtype=torch.int64, device=dev)
    ops.layer_fwd(X, W, b, P, Cin, Cout, in_scale=sc, in_shift=sh, in_relu=True, Y=Y, stat_partial=part,
                  precision=1, tc_ws=ws, tc_packed=True, debug_flags=0, debug_clocks=clk, **kw)
    torch.cuda.synchronize()
    c = clk.view(148, 17, 8).double().cpu().numpy()
    tot = c[:, :, 7].mean()
    def pct(role, names):
        m = c[:, role, :].reshape(-1, 8).mean(0)
        return {n: round(100 * m[i] / tot, 1) for i, n in enumerate(names) if n}
    res["clk_total_kcycles"] = round(tot / 1e3, 1)
    res["clk_epilogue_%"] = pct(slice(0, 8), ["wait_tfull", "tmem_ld+bias", "stage_sts", "y_store", "stats", "tail"])
    res["clk_mma_%"] = pct(slice(8, 9), ["wait_tempty", "wait_full", "issue"])
    res["clk_producer_%"] = pct(slice(9, 17), ["wait_empty", "tma+convert+sts", "fetch+fence+arrive", "loop"])
    flops = 2.0 * P * Cin * Cout
    res["mma_floor_us_at_1965MHz"] = round(3 * flops / (148 * 2048 * 2 * 1.965e9) * 1e6, 1)
    out["%dx%d->%d%s" % (P, Cin, Cout, " g16 noY" if group else "")] = res
print(json.dumps(out, indent=1))
