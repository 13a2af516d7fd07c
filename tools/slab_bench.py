#!/usr/bin/env python3
"""BASELINE config C5b: ONE large grid (default 1024^3, rand-flexi + learnable deform, fp32) sharded
into slabs along tensor dim 0 across the ranks of a torchrun job, DiffMC and DiffDMC forward+backward
with halo exchange and global-id stitching over NCCL (diso_b200/parallel.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/slab_bench.py [--size 1024]

Every rank generates only ITS slab, on the device, from per-layer seeds (no rank ever holds the whole
grid).  Timing: CUDA events around K steps, barrier + synchronize on both sides, max over ranks.
Rank 0 prints one JSON line (Gvoxel/s counts the voxels of the whole grid, 2 extractor passes per step)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from diso_b200 import parallel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
a = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = a.size
xa, xb = parallel.plan_slabs(n, world)[rank]


def layer(x, shape, lo, hi):
    g = torch.Generator(device=dev).manual_seed(1000003 * x + 17)
    return torch.rand(shape, generator=g, device=dev) * (hi - lo) + lo


sdf = torch.stack([layer(x, (n, n), -0.1, 0.9) for x in range(xa, xb)]).requires_grad_(True)              # rand-flexi: U(0,1) - 0.1
deform = torch.stack([0.5 * torch.tanh(layer(x + n, (n, n, 3), 0.0, 1.0)) for x in range(xa, xb)]).requires_grad_(True)


def step():
    out = {}
    for alg in ("mc", "dmc"):
        sdf.grad = deform.grad = None
        verts, faces, info = parallel.extract_slab(alg, sdf, deform, (xa, xb), n, 0.0, True)
        (verts * 0.5).sum().backward()
        out[alg] = (info["n_verts_total"], info["n_faces_total"])
        del verts, faces
    return out


def sync():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()


for _ in range(a.warmup):
    out = step()
sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    out = step()
e1.record()
sync()
t = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
mem = torch.tensor([torch.cuda.max_memory_allocated() / 2 ** 30], dtype=torch.float64, device=dev)
dist.all_reduce(mem, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t.item())
    print(json.dumps({"workload": "C5b: one %d^3 grid (rand-flexi + deform, fp32) in %d slabs along dim 0, DiffMC + DiffDMC fwd+bwd per step" % (n, world),
                      "n_gpus": world, "ms_per_step": ms, "value": 2 * n ** 3 / (ms * 1e-3) / 1e9, "unit": "Gvoxel/s",
                      "mesh": {k: dict(verts=v[0], faces=v[1]) for k, v in out.items()}, "max_gib_per_gpu": round(float(mem.item()), 1),
                      "steps": a.steps, "warmup": a.warmup}), flush=True)
dist.destroy_process_group()
