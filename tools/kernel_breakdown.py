#!/usr/bin/env python3
"""Per-kernel device times of one forward+backward: python tools/kernel_breakdown.py sphere|flexi|sparse|dense SIZE [mc|dmc] [nodeform]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diso_b200
from diso_b200 import _lib, synthetic as syn
kind, n = sys.argv[1], int(sys.argv[2])
alg = sys.argv[3] if len(sys.argv) > 3 else "mc"
use_def = not (len(sys.argv) > 4 and sys.argv[4] == "nodeform")
sdf = (syn.sphere_sdf(n) if kind == "sphere" else syn.random_sdf(n, kind, 0)).cuda().requires_grad_(True)
deform = syn.random_deform(n, 1).cuda().requires_grad_(True) if use_def else None
m = diso_b200.DiffMC() if alg == "mc" else diso_b200.DiffDMC()
kw = {} if alg == "mc" else dict(return_quads=True)
for _ in range(3):
    v, f = m(sdf, deform, **kw); v.sum().backward()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with _lib.kernel_profile() as prof:
    e0.record()
    for _ in range(5):
        sdf.grad = None
        v, f = m(sdf, deform, **kw); v.sum().backward()
    e1.record(); torch.cuda.synchronize()
print("%s %d^3 %s deform=%s: %d verts %d faces, %.3f ms per fwd+bwd" % (kind, n, alg, use_def, v.shape[0], f.shape[0], e0.elapsed_time(e1) / 5))
for k, x in sorted(prof.times.items(), key=lambda kv: -sum(kv[1])):
    print("   %-20s %.3f ms" % (k, sum(x) / 5))
