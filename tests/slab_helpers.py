"""Helpers for the slab-sharding tests: an oracle-backed local extractor (CPU) with the same
interface as diso_b200.parallel.cuda_extractor, and the per-rank worker body."""
import os
import re

import numpy as np
import torch
from torch.autograd import Function

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tables():
    src = open(os.path.join(ROOT, "oracle", "diso_tables.h")).read()
    out = {}
    for m in re.finditer(r"(\w+)\[(\d+)\] = \{(.*?)\};", src, re.S):
        out[m.group(1)] = np.array([int(x) for x in re.findall(r"-?\d+", m.group(3))])
    return out


def oracle_extractor(alg):
    from oracle import diso_oracle as O
    T = _tables()
    per_code = (np.diff(T["T_MC_FIRST"]) // 3) if alg == "mc" else np.diff(T["T_PATCH_FIRST"])

    class Fn(Function):
        @staticmethod
        def forward(ctx, sdf, deform, iso):
            s = sdf.detach().numpy()
            d = deform.detach().numpy() if deform is not None else None
            g, dp = O.pad_inputs(s, d, iso)
            r = O.raw_forward(alg, g, dp, iso)
            ctx.pack = (g, dp, iso, deform is not None)
            ctx.raw = r
            verts = torch.from_numpy(r["verts"] - s.dtype.type(1))
            faces = torch.from_numpy(r["faces"].astype(np.int64))
            ctx.mark_non_differentiable(faces)
            return verts, faces

        @staticmethod
        def backward(ctx, adj_verts, _):
            g, dp, iso, has_def = ctx.pack
            ag, ad = O.raw_backward(alg, g, dp, iso, adj_verts.contiguous().numpy(), "reference")
            ag = torch.from_numpy(np.ascontiguousarray(ag[1:-1, 1:-1, 1:-1]))
            ad = torch.from_numpy(np.ascontiguousarray(ad[1:-1, 1:-1, 1:-1])) if has_def else None
            return ag, ad, None

    def run(sdf_ext, deform_ext, iso):
        fn_out = Fn.apply(sdf_ext, deform_ext, iso)
        verts, faces = fn_out
        # per padded layer prefix sums, computed independently from the sign field
        s = sdf_ext.detach().numpy()
        g, _ = O.pad_inputs(s, None, iso)
        b = g >= g.dtype.type(iso)
        PX = b.shape[0]
        edges = np.zeros(PX, np.int64)
        edges[:-1] += (b[:-1] != b[1:]).reshape(PX - 1, -1).sum(1)
        edges += (b[:, :-1] != b[:, 1:]).reshape(PX, -1).sum(1)
        edges += (b[:, :, :-1] != b[:, :, 1:]).reshape(PX, -1).sum(1)
        e_pre = np.concatenate([[0], np.cumsum(edges)])
        raw = O.raw_forward(alg, g, None, iso)
        layer = raw["used_index"].astype(np.int64) // (b.shape[1] * b.shape[2])
        f_cnt = np.bincount(layer, weights=per_code[raw["used_code"]], minlength=PX).astype(np.int64)
        f_pre = np.concatenate([[0], np.cumsum(f_cnt)])
        return verts, faces, torch.from_numpy(e_pre), torch.from_numpy(f_pre)
    return run


def worker(rank, world, port, alg, sdf, deform, iso, normalize, use_cuda, out_dir, field=False):
    """Body of one rank: shard, extract, backward with a fixed dL/dverts, dump results.
    field=True: the rank keeps its slab as a parallel.SlabField (extended leaf, in-place halo refresh)."""
    import torch.distributed as dist
    from diso_b200 import parallel
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        X = sdf.shape[0]
        a, b = parallel.plan_slabs(X, world)[rank]
        dev = "cuda:0" if use_cuda else "cpu"
        s_own = sdf[a:b].clone().to(dev).requires_grad_(True)
        d_own = deform[a:b].clone().to(dev).requires_grad_(True) if deform is not None else None
        ext = None if use_cuda else oracle_extractor(alg)
        if field:
            sf = parallel.SlabField(s_own.detach(), rank, world)
            sf.ext.requires_grad_(True)
            df = None
            if d_own is not None:
                df = parallel.SlabField(d_own.detach(), rank, world)
                df.ext.requires_grad_(True)
            verts, faces, info = parallel.extract_slab_ext(alg, sf, df, (a, b), X, iso, normalize)
        else:
            verts, faces, info = parallel.extract_slab(alg, s_own, d_own, (a, b), X, iso, normalize, extractor=ext)
        if verts.shape[0] or verts.requires_grad:   # (an empty but attached result keeps the halo backward symmetric)
            i = torch.arange(info["vert_offset"] * 3, (info["vert_offset"] + verts.shape[0]) * 3, dtype=torch.float64).reshape(-1, 3)
            w = torch.cos(i * 0.6180339887 + 0.25).to(verts.dtype).to(dev)
            (verts * w).sum().backward()
        elif not field:
            (s_own.sum() * 0).backward()
        if field:   # gradients of the rank's own layers (halo layers of the extended leaf carry zeros after the exchange)
            z = torch.zeros_like(sf.ext)
            ge = sf.ext.grad if sf.ext.grad is not None else z
            assert float(ge[: sf.n_lo].abs().sum()) == 0 and float(ge[sf.n_lo + sf.n:].abs().sum()) == 0
            s_own.grad = ge[sf.n_lo: sf.n_lo + sf.n].clone()
            if df is not None:
                gd = df.ext.grad if df.ext.grad is not None else torch.zeros_like(df.ext)
                d_own.grad = gd[df.n_lo: df.n_lo + df.n].clone()
        torch.save(dict(verts=verts.detach().cpu(), faces=faces.cpu(), info=info,
                        gsdf=s_own.grad.cpu() if s_own.grad is not None else torch.zeros_like(s_own).cpu(),
                        gdef=(d_own.grad.cpu() if d_own.grad is not None else torch.zeros_like(d_own).cpu()) if d_own is not None else None),
                   os.path.join(out_dir, "rank%d.pt" % rank))
    finally:
        dist.destroy_process_group()


def field_worker(rank, world, port, full, out_dir):
    """SlabField mechanics on the CPU (gloo): in-place halo refresh and the gradient return of the halo layers."""
    import torch.distributed as dist
    from diso_b200 import parallel
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        X = full.shape[0]
        a, b = parallel.plan_slabs(X, world)[rank]
        f = parallel.SlabField(full[a:b].clone(), rank, world)
        f.ext.requires_grad_(True)
        ext = f.synced()
        lo = a - f.n_lo
        assert torch.equal(ext.detach(), full[lo: b + f.n_hi]), "halo layers differ from the neighbours' layers"
        # a loss that touches every layer of the extended slab with a global-position-dependent weight
        w = torch.arange(lo, b + f.n_hi, dtype=full.dtype).view(-1, *([1] * (full.dim() - 1))) + 1.0
        (ext * ext * w).sum().backward()
        g = f.ext.grad
        assert float(g[: f.n_lo].abs().sum()) == 0 and float(g[f.n_lo + f.n:].abs().sum()) == 0, "halo gradients must be returned, not kept"
        torch.save(dict(a=a, b=b, grad=g[f.n_lo: f.n_lo + f.n].clone()), os.path.join(out_dir, "field%d.pt" % rank))
    finally:
        dist.destroy_process_group()
