#!/usr/bin/env python3
"""cProfile of the host side of a small extraction (sphere 64^3, DiffMC forward+backward): which Python / ctypes / torch calls
make up the ~0.2 ms that the five ~10 us kernels do not explain.  python tools/host_profile.py [iterations]"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diso_b200
from diso_b200 import synthetic as syn

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
s = syn.sphere_sdf(64).to("cuda:0").requires_grad_(True)
m = diso_b200.DiffMC()


def fwd_only():
    for _ in range(N):
        v, f = m(s)


def fwd_bwd():
    for _ in range(N):
        s.grad = None
        v, f = m(s)
        v.sum().backward()


for fn in (fwd_only, fwd_bwd):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("== %s: %.1f us per iteration (unprofiled)" % (fn.__name__, (t1 - t0) / N * 1e6))
    pr = cProfile.Profile()
    pr.enable(); fn(); torch.cuda.synchronize(); pr.disable()
    st = pstats.Stats(pr, stream=sys.stdout)
    st.sort_stats("tottime").print_stats(22)
