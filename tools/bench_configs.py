#!/usr/bin/env python3
"""Times the BASELINE.md configs C1-C4 (forward+backward, median of N iterations, CUDA events) for
this library and for the unmodified reference CUDA build (baseline/_ref) on the same GPU.
Prints a markdown table; used for BASELINE.md section 4."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import diso_b200  # noqa: E402
from diso_b200 import synthetic as syn  # noqa: E402
from tests.refload import load_reference  # noqa: E402  (tools may use the test helpers; bench.py does not)

dev = "cuda:0"
ref = load_reference()


def timeit(mod, sdf, deform, kw, iters, warm=3):
    s = sdf.to(dev).requires_grad_(True)
    d = deform.to(dev).requires_grad_(True) if deform is not None else None
    v, f = mod(s, d, **kw)
    gen = torch.Generator().manual_seed(7)
    w = torch.rand(v.shape, generator=gen).to(v.dtype).to(dev)
    ts = []
    for i in range(warm + iters):
        s.grad = None
        if d is not None:
            d.grad = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        v, f = mod(s, d, **kw)
        (v * w).sum().backward()
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], v.shape[0], f.shape[0]


CFG = [
    ("C1 sphere 64^3, DiffMC fp32, no deform", "mc", torch.float32, lambda dt: syn.sphere_sdf(64, dtype=dt), False, {}, 100),
    ("C2 round cube 128^3, DiffDMC fp32 + deform, quads", "dmc", torch.float32, lambda dt: syn.round_cube_sdf(128, dtype=dt), True, dict(return_quads=True), 100),
    ("C2' round cube 128^3, DiffDMC fp32 + deform, triangles", "dmc", torch.float32, lambda dt: syn.round_cube_sdf(128, dtype=dt), True, dict(return_quads=False), 50),
]
for kind in ("flexi", "sparse", "dense"):
    for alg in ("mc", "dmc"):
        CFG.append(("C3 rand-%s 256^3, %s fp32" % (kind, "DiffMC" if alg == "mc" else "DiffDMC quads"), alg, torch.float32,
                    (lambda k: (lambda dt: syn.random_sdf(256, k, 0, dt)))(kind), False, {} if alg == "mc" else dict(return_quads=True), 30))
for dt_ in (torch.float32, torch.float64):
    for alg in ("mc", "dmc"):
        CFG.append(("C4 rand-flexi 512^3 + deform, %s %s" % ("DiffMC" if alg == "mc" else "DiffDMC quads", "fp32" if dt_ == torch.float32 else "fp64"),
                    alg, dt_, lambda dt: syn.random_sdf(512, "flexi", 0, dt), True, {} if alg == "mc" else dict(return_quads=True), 10))
CFG.append(("sphere 512^3, DiffMC fp32 + deform (sparse surface)", "mc", torch.float32, lambda dt: syn.sphere_sdf(512, dtype=dt), True, {}, 20))

print("| config | verts | faces | ours ms | reference CUDA ms | speed-up | ours Gvoxel/s |")
print("|---|---|---|---|---|---|---|")
for name, alg, dt, mk, use_def, kw, iters in CFG:
    sdf = mk(dt)
    deform = syn.random_deform(tuple(sdf.shape), 1, dt) if use_def else None
    ours = diso_b200.DiffMC(dt) if alg == "mc" else diso_b200.DiffDMC(dt)
    t, nv, nf = timeit(ours, sdf, deform, kw, iters)
    tr = None
    if ref is not None:
        try:
            theirs = ref.DiffMC(dt) if alg == "mc" else ref.DiffDMC(dt)
            tr, _, _ = timeit(theirs, sdf, deform, kw, max(3, iters // 3))
        except Exception as ex:  # e.g. out of memory in the reference
            tr = None
    print("| %s | %d | %d | %.3f | %s | %s | %.2f |" % (name, nv, nf, t, "%.3f" % tr if tr else "n/a", "%.1fx" % (tr / t) if tr else "n/a",
                                                          sdf.numel() / t / 1e6), flush=True)
    del sdf, deform
    torch.cuda.empty_cache()
