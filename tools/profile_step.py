#!/usr/bin/env python3
"""Minimal driver for ncu: N steps of DiffMC + DiffDMC forward+backward on a random-init grid
(the bench.py workload without the bench machinery)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import diso_b200  # noqa: E402
from diso_b200 import synthetic as syn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--kind", default="flexi")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--alg", default="both")
a = ap.parse_args()
dev = "cuda:0"
sdf = syn.random_sdf(a.size, a.kind, 0).to(dev).requires_grad_(True)
deform = syn.random_deform(a.size, 1000).to(dev).requires_grad_(True)
mods = []
if a.alg in ("both", "mc"):
    mods.append((diso_b200.DiffMC(), {}))
if a.alg in ("both", "dmc"):
    mods.append((diso_b200.DiffDMC(), dict(return_quads=True)))
for _ in range(a.steps):
    for m, kw in mods:
        sdf.grad = None
        deform.grad = None
        v, f = m(sdf, deform, **kw)
        v.sum().backward()
torch.cuda.synchronize()
print("done", v.shape, f.shape)
