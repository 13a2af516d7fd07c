#!/usr/bin/env python3
"""Which summation order do torch's CUDA reductions use for rows of 3?  (GPU probe, prints match counts.)

The reference's quad split (diso/__init__.py:118-147) uses F.normalize(v, dim=-1) and (a * b).sum(-1) on [Q,3]
tensors; the diagonal choice depends on the last bit of those reductions.  This probe evaluates candidate orders
with elementwise torch ops (one rounding per op) and reports which one reproduces torch bit for bit."""
import itertools
import sys

import torch
import torch.nn.functional as F


def fma32(a, b, c):   # RN32(a*b + c): the product is exact in fp64
    return (a.double() * b.double() + c.double()).float()


def main():
    dev = "cuda:0"
    g = torch.Generator(device="cpu").manual_seed(0)
    for dt in (torch.float32, torch.float64):
        v = (torch.rand((1 << 22, 3), generator=g, dtype=torch.float64) - 0.5).to(dt).to(dev)
        w = (torch.rand((1 << 22, 3), generator=g, dtype=torch.float64) - 0.5).to(dt).to(dev)
        x, y, z = v.unbind(-1)
        n_ref = torch.linalg.vector_norm(v, dim=-1)
        sq = {"x": x * x, "y": y * y, "z": z * z}
        print("dtype", dt)
        for a, b, c in itertools.permutations("xyz"):
            cand = torch.sqrt((sq[a] + sq[b]) + sq[c])
            print("  norm sqrt((%s2+%s2)+%s2): mismatches %d" % (a, b, c, int((cand != n_ref).sum())))
        if dt == torch.float32:
            comp = {"x": x, "y": y, "z": z}
            for a, b, c in itertools.permutations("xyz"):
                cand = torch.sqrt(fma32(comp[c], comp[c], fma32(comp[b], comp[b], comp[a] * comp[a])))
                print("  norm sqrt(fma(%s,%s,fma(%s,%s,%s2))): mismatches %d" % (c, c, b, b, a, int((cand != n_ref).sum())))
        # F.normalize = v / max(norm, eps)
        nn = F.normalize(v, dim=-1)
        cand = v / n_ref.clamp_min(1e-12)[:, None]
        print("  normalize == v / clamp_min(vector_norm): mismatches", int((cand != nn).sum()))
        p = v * w
        s_ref = p.sum(-1)
        px, py, pz = p.unbind(-1)
        pp = {"x": px, "y": py, "z": pz}
        for a, b, c in itertools.permutations("xyz"):
            cand = (pp[a] + pp[b]) + pp[c]
            print("  sum (%s+%s)+%s: mismatches %d" % (a, b, c, int((cand != s_ref).sum())))
        # torch.max over a stacked last dim: any NaN / order subtleties?  values only, order-free.
    return 0


if __name__ == "__main__":
    sys.exit(main())
