#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/slab_nccl_check.py [--size 256]: slab-sharded extraction of one
grid over NCCL, checked on rank 0 against the single-GPU operator on the whole grid; prints timings."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import diso_b200  # noqa: E402
from diso_b200 import parallel, synthetic as syn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
a = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = a.size
sdf = syn.random_sdf(n, "flexi", 0)
deform = syn.random_deform(n, 1000)
xa, xb = parallel.plan_slabs(n, world)[rank]
ok = True
for alg in ("mc", "dmc"):
    for api in ("own", "field"):
        s_own = sdf[xa:xb].to(dev).requires_grad_(True)
        d_own = deform[xa:xb].to(dev).requires_grad_(True)
        if api == "field":      # extended leaves, halos refreshed in place (parallel.SlabField)
            sf, df = parallel.SlabField(s_own.detach(), rank, world), parallel.SlabField(d_own.detach(), rank, world)
            sf.ext.requires_grad_(True); df.ext.requires_grad_(True)
        for it in range(3):
            torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
            if api == "field":
                sf.ext.grad = df.ext.grad = None
                verts, faces, info = parallel.extract_slab_ext(alg, sf, df, (xa, xb), n, 0.0, True)
            else:
                s_own.grad = d_own.grad = None
                verts, faces, info = parallel.extract_slab(alg, s_own, d_own, (xa, xb), n, 0.0, True)
            (verts * 1.0).sum().backward()
            torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
        grad = sf.ext.grad[sf.n_lo: sf.n_lo + sf.n] if api == "field" else s_own.grad
        # gather everything on rank 0 for the check
        parts_v = [None] * world; parts_f = [None] * world; parts_g = [None] * world
        dist.all_gather_object(parts_v, verts.detach().cpu()); dist.all_gather_object(parts_f, faces.cpu())
        dist.all_gather_object(parts_g, grad.cpu())
        if rank == 0:
            s = sdf.to(dev).requires_grad_(True); d = deform.to(dev).requires_grad_(True)
            m = diso_b200.DiffMC() if alg == "mc" else diso_b200.DiffDMC()
            kw = {} if alg == "mc" else dict(return_quads=True)
            torch.cuda.synchronize(); t2 = time.perf_counter()
            ev, ef = m(s, d, **kw); ev.sum().backward(); torch.cuda.synchronize(); t3 = time.perf_counter()
            V, F, G = torch.cat(parts_v), torch.cat(parts_f), torch.cat(parts_g)
            same_f = torch.equal(F, ef.cpu())
            same_v = torch.equal(V, ev.detach().cpu())       # the frame makes slab vertices bit-identical
            dg = float((G - s.grad.cpu()).abs().max() / max(1.0, float(s.grad.abs().max())))
            ok &= same_f and same_v and dg < 1e-5
            print("%s %d^3 world=%d api=%s: faces_equal=%s verts_bit_identical=%s rel|dgrad|=%.2e  sharded %.1f ms vs single-GPU %.1f ms (cold)"
                  % (alg, n, world, api, same_f, same_v, dg, (t1 - t0) * 1e3, (t3 - t2) * 1e3), flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
