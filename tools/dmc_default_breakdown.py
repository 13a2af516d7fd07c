#!/usr/bin/env python3
"""Per-kernel device times of DiffDMC's DEFAULT call (return_quads=False) forward+backward: python tools/dmc_default_breakdown.py [SIZE]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diso_b200
from diso_b200 import _lib, synthetic as syn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sdf = syn.random_sdf(n, "flexi", 0).cuda().requires_grad_(True)
deform = syn.random_deform(n, 1).cuda().requires_grad_(True)
m = diso_b200.DiffDMC()
for _ in range(3):
    v, f = m(sdf, deform); v.sum().backward()
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
with _lib.kernel_profile() as prof:
    fwd = 0.0
    for _ in range(5):
        sdf.grad = None
        e0.record(); v, f = m(sdf, deform); e1.record(); v.sum().backward(); e2.record(); torch.cuda.synchronize()
        fwd += e0.elapsed_time(e1)
print("flexi %d^3 DiffDMC default: %d verts %d tris, forward %.3f ms (flags fused: %s)" % (n, v.shape[0], f.shape[0], fwd / 5, not os.environ.get("DISO_B200_NO_QUAD_FLAGS")))
for k, x in sorted(prof.times.items(), key=lambda kv: -sum(kv[1])):
    print("   %-20s %.3f ms" % (k, sum(x) / 5))
