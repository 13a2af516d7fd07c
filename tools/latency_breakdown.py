#!/usr/bin/env python3
"""Host-side latency breakdown of one small extraction (where do the ~0.3 ms of a 64^3 call go?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diso_b200
from diso_b200 import _lib, synthetic as syn

dev = "cuda:0"
for name, sdf in (("sphere64", syn.sphere_sdf(64)), ("roundcube128", syn.round_cube_sdf(128))):
    s = sdf.to(dev).requires_grad_(True)
    m = diso_b200.DiffMC()
    for _ in range(20):
        v, f = m(s); v.sum().backward()
    torch.cuda.synchronize()
    N = 200
    t0 = time.perf_counter()
    for _ in range(N):
        st, c = diso_b200._count(_lib.ALG_MC, s.detach(), 0.0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    for _ in range(N):
        v, f = m(s)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    for _ in range(N):
        s.grad = None
        v, f = m(s); v.sum().backward()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    with torch.no_grad():
        for _ in range(N):
            v, f = m(s)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    print("%s: count+sync %.1f us | forward %.1f us | forward(no_grad) %.1f us | fwd+bwd %.1f us" % (
        name, (t1 - t0) / N * 1e6, (t2 - t1) / N * 1e6, (t4 - t3) / N * 1e6, (t3 - t2) / N * 1e6))
    with _lib.kernel_profile() as prof:
        s.grad = None
        v, f = m(s); v.sum().backward()
    print("   kernel device times (us):", {k: round(1e3 * sum(x), 1) for k, x in prof.times.items()})
