#!/usr/bin/env python
"""One tcgen05 layer shape, a few launches: the target of `ncu --set full --import-source on` (tools/ncu_tc_one.sh)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from usip_b200 import ops
P, Cin, Cout = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (131072, 256, 256)))
group = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dev = torch.device("cuda:0")
X = torch.randn(P, Cin, device=dev); W = torch.randn(Cout, Cin, device=dev) / Cin ** 0.5
b = torch.randn(Cout, device=dev); sc = torch.rand(Cin, device=dev) + 0.5; sh = torch.randn(Cin, device=dev)
Y = None if group else torch.empty(P, Cout, device=dev)
part = torch.empty(ops.stat_slots(P, Cout, 1, group, bool(group)), 2, Cout, device=dev)
ws = torch.empty(2 * Cin * Cout, device=dev)
kw = dict(gmax=torch.empty(P // group, Cout, device=dev), gmin=torch.empty(P // group, Cout, device=dev), group=group) if group else {}
for i in range(4):
    ops.layer_fwd(X, W, b, P, Cin, Cout, in_scale=sc, in_shift=sh, in_relu=True, Y=Y, stat_partial=part, precision=1,
                  tc_ws=ws, tc_packed=i > 0, **kw)
torch.cuda.synchronize()
