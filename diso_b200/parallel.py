"""Slab sharding of ONE large grid across ranks (BASELINE config C5b).

The reference has no multi-device support at all.  Shapes in a batch are independent and need
no code (one extractor call per GPU, see bench.py); this module covers the other case: a single
grid too large (or too slow) for one GPU, split into slabs along tensor dim 0 -- the slowest
axis, so every slab is contiguous in memory AND, because every ordering of the reference is
ascending in the padded linear index, the global output is simply the rank-order concatenation
of what the ranks own.

Ownership (padded x-layers xp in [0, X+2)): rank r owns the layers [A_r, B_r) -- its real layers
plus, for the first / last rank, the global pad layers.  It owns the crossing edges (MC
vertices / DMC quads) whose start point lies in those layers and the cells (MC triangles / DMC
dual vertices) whose origin does.

Each rank extracts its slab extended by a 2-layer halo on both sides with the ordinary
single-GPU operator.  Two layers are exactly what keeps everything a rank OWNS, and every id it
REFERENCES in a neighbour's range, unaffected by the virtual iso+1 pad the local extraction puts
at the cut: triangles reference vertices on layer B_r (their rank inside the neighbour's
numbering needs layer B_r's own x-edges, i.e. the signs of layer B_r+1); DMC quads reference dual
vertices of cell layer A_r-1, whose ambiguity flip looks at cell layer A_r-2.
Data-path traffic over NVLink: the halo layers (2*Y*Z values per neighbour, forward and backward)
and one all_gather of three integers per rank -- never dense data.

  * owned items are contiguous ranges of the local output: [prefix[A_r-lo] : prefix[B_r-lo]) with
    the per-layer prefix sums read from the extractor's state (diso_b200.layer_prefixes);
  * local ids -> global ids by adding per-range constants (offsets from the all_gather);
  * gradients: the halo exchange is an autograd Function; its backward sends the gradients of the
    halo layers back to their owners, which add them to their own.

Vertices are computed in the slab's local frame and shifted by the slab origin afterwards, so
they agree with a single-GPU run to rounding (<= 1 ulp of the coordinate), not bit-for-bit;
connectivity is identical.
"""
import torch
import torch.distributed as dist
from torch.autograd import Function

HALO = 2


def plan_slabs(X, world):
    """Contiguous, balanced x-ranges [(a_0,b_0), ...]; every slab gets at least HALO layers."""
    if X < HALO * world:
        raise ValueError("grid dim 0 (%d) too small for %d slabs of >= %d layers" % (X, world, HALO))
    base, rem = divmod(X, world)
    out, a = [], 0
    for r in range(world):
        b = a + base + (1 if r < rem else 0)
        out.append((a, b))
        a = b
    return out


def _p2p(ops_spec, group):
    """ops_spec: list of (kind, tensor, peer).  gloo cannot move CUDA tensors: stage through the host."""
    if not ops_spec:
        return
    staged, ops = [], []
    gloo = dist.get_backend(group) == "gloo"
    for kind, t, peer in ops_spec:
        buf = t
        if gloo and t.is_cuda:
            buf = t.cpu() if kind == "send" else torch.empty(t.shape, dtype=t.dtype)
            staged.append((kind, t, buf))
        ops.append(dist.P2POp(dist.isend if kind == "send" else dist.irecv, buf.contiguous() if kind == "send" else buf, peer, group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    for kind, t, buf in staged:
        if kind == "recv":
            t.copy_(buf)


class _HaloExchange(Function):
    """own [n,...] -> [halo_lo ; own ; halo_hi] with HALO layers from each existing neighbour."""

    @staticmethod
    def forward(ctx, own, rank, world, group):
        ctx.rank, ctx.world, ctx.group = rank, world, group
        lo = own.new_empty((HALO,) + own.shape[1:]) if rank > 0 else None
        hi = own.new_empty((HALO,) + own.shape[1:]) if rank < world - 1 else None
        spec = []
        if rank > 0:
            spec += [("send", own[:HALO].contiguous(), rank - 1), ("recv", lo, rank - 1)]
        if rank < world - 1:
            spec += [("send", own[-HALO:].contiguous(), rank + 1), ("recv", hi, rank + 1)]
        _p2p(spec, group)
        parts = ([lo] if lo is not None else []) + [own] + ([hi] if hi is not None else [])
        return torch.cat(parts, 0)

    @staticmethod
    def backward(ctx, g):
        rank, world, group = ctx.rank, ctx.world, ctx.group
        n_lo = HALO if rank > 0 else 0
        n_hi = HALO if rank < world - 1 else 0
        g_own = g[n_lo: g.shape[0] - n_hi].clone()
        from_prev = g_own.new_empty((HALO,) + g.shape[1:]) if rank > 0 else None
        from_next = g_own.new_empty((HALO,) + g.shape[1:]) if rank < world - 1 else None
        spec = []
        if rank > 0:        # my lo halo belongs to prev's last layers; prev's hi halo are my first layers
            spec += [("send", g[:HALO].contiguous(), rank - 1), ("recv", from_prev, rank - 1)]
        if rank < world - 1:
            spec += [("send", g[g.shape[0] - HALO:].contiguous(), rank + 1), ("recv", from_next, rank + 1)]
        _p2p(spec, group)
        if from_prev is not None:
            g_own[:HALO] += from_prev
        if from_next is not None:
            g_own[-HALO:] += from_next
        return g_own, None, None, None


def cuda_extractor(alg, dtype=torch.float32, grad_mode="reference"):
    """The product extractor for slabs: the ordinary single-GPU operator + its per-layer prefixes."""
    import diso_b200
    from diso_b200 import _lib
    alg_id = {"mc": _lib.ALG_MC, "dmc": _lib.ALG_DMC}[alg]
    gm = {"reference": _lib.GRAD_REFERENCE, "exact": _lib.GRAD_EXACT}[grad_mode]

    def run(sdf_ext, deform_ext, isovalue):
        verts, faces, state = diso_b200._run(alg_id, dtype, gm, sdf_ext, deform_ext, isovalue, False,
                                             want_state=True, slab_mode=True)
        if state is None:
            n = sdf_ext.shape[0] + 3
            z = torch.zeros(n, dtype=torch.int64)
            return verts, faces.long(), z, z.clone()
        e, f = diso_b200.layer_prefixes(alg_id, state, tuple(sdf_ext.shape))
        return verts, faces, e, f
    return run


def _gather_counts(mine, group, world):
    if world == 1:
        return mine.cpu()[None]
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    if dist.get_backend(group) == "gloo" and mine.is_cuda:
        cpu = [g.cpu() for g in gathered]
        dist.all_gather(cpu, mine.cpu(), group=group)
        return torch.stack(cpu)
    dist.all_gather(gathered, mine, group=group)
    return torch.stack(gathered).cpu()


def _extract_slab_fused(alg, sdf_own, deform_own, a, b, X, isovalue, normalize, group, rank, world, grad_mode="reference"):
    """The product path of :func:`extract_slab`: count -> exchange the per-rank totals -> emit with a
    frame (include/diso_b200.h: diso_b200_frame), so the kernels write global-frame vertices (bit-identical
    to a single extraction of the whole grid) and global vertex ids directly; what a rank owns are
    contiguous slices (views) of the emitted arrays.  No elementwise pass touches the outputs."""
    import diso_b200
    from diso_b200 import _lib
    alg_id = {"mc": _lib.ALG_MC, "dmc": _lib.ALG_DMC}[alg]
    gm = {"reference": _lib.GRAD_REFERENCE, "exact": _lib.GRAD_EXACT}[grad_mode]
    dev, dt = sdf_own.device, sdf_own.dtype
    k = 3 if alg == "mc" else 4
    diso_b200._check_inputs(sdf_own, deform_own, dt)
    sdf_ext = _HaloExchange.apply(sdf_own, rank, world, group) if world > 1 else sdf_own
    def_ext = None
    if deform_own is not None:
        def_ext = _HaloExchange.apply(deform_own, rank, world, group) if world > 1 else deform_own
    lo = a - (HALO if rank > 0 else 0)                      # global x of the first local layer
    A = a + 1 if rank > 0 else 0                            # owned padded layers [A, B), global padded coords
    B = b + 1 if rank < world - 1 else X + 2
    lA, lB = A - lo, B - lo
    with torch.cuda.device(dev):
        g = sdf_ext.contiguous()
        d = def_ext.contiguous() if def_ext is not None else None
        with torch.no_grad():
            state, counts = diso_b200._count(alg_id, g, isovalue)
        e0 = e1 = f0 = f1 = 0
        if counts[_lib.CNT_EDGES] > 0:
            e_pre, f_pre = diso_b200.layer_prefixes(alg_id, state, tuple(g.shape))
            e0, e1 = int(e_pre[lA]), int(e_pre[lB])       # crossing edges: MC vertices / DMC quads
            f0, f1 = int(f_pre[lA]), int(f_pre[lB])       # MC triangles / DMC dual vertices
        v0, v1, q0, q1 = (e0, e1, f0, f1) if alg == "mc" else (f0, f1, e0, e1)   # owned vertex / face ranges (local ids)
        # "some value > iso" over the extended slab; the halo layers are other ranks' layers, so the OR over the
        # ranks is the global test of the reference's early-out (diso/__init__.py:49)
        mine = torch.tensor([v1 - v0, q1 - q0, counts[_lib.CNT_ANY_GT]], dtype=torch.int64, device=dev)
        allc = _gather_counts(mine, group, world)
        v_off = int(allc[:rank, 0].sum())
        info = dict(vert_offset=v_off, face_offset=int(allc[:rank, 1].sum()), n_verts_total=int(allc[:, 0].sum()),
                    n_faces_total=int(allc[:, 1].sum()), owned_layers=(A, B))
        if info["n_verts_total"] == 0 or int(allc[:, 2].sum()) == 0:
            info.update(n_verts_total=0, n_faces_total=0, vert_offset=0, face_offset=0)
            return torch.zeros((0, 3), dtype=dt, device=dev), torch.zeros((0, k), dtype=torch.int64, device=dev), info
        if v1 == v0 and q1 == q0:
            # nothing owned here, but the neighbours' backward still exchanges halo gradients with this rank:
            # keep the (empty) result attached to the halo exchange so its backward runs
            verts = sdf_ext.flatten()[:0].reshape(0, 1).expand(0, 3)
            if def_ext is not None:
                verts = verts + def_ext.flatten()[:0].reshape(0, 1)
            return verts, torch.zeros((0, k), dtype=torch.int64, device=dev), info
        verts_l, faces_l = diso_b200._Extract.apply(g, d, alg_id, float(isovalue), bool(normalize), gm, state, counts,
                                                    (lo, X, v_off - v0))
        return verts_l[v0:v1], faces_l[q0:q1], info


def extract_slab(alg, sdf_own, deform_own, x_range, X, isovalue=0.0, normalize=True, group=None, extractor=None):
    """Slab-sharded extraction.  Every rank of `group` calls this with ITS slab
    ``sdf_own = sdf[a:b]`` (and ``deform_own = deform[a:b]`` or None), ``x_range = (a, b)`` from
    :func:`plan_slabs` and the global ``X``.

    Returns ``(verts, faces, info)``: the vertices this rank owns ([n,3], global frame, divided by
    the global ``dims-1`` if ``normalize``), the faces it owns with GLOBAL vertex ids (MC [F,3]
    triangles, DMC [Q,4] quads; int64) and ``info`` with the global offsets / totals.  The global
    mesh is the rank-order concatenation.  Gradients flow to ``sdf_own`` / ``deform_own`` of every
    rank (halo contributions are exchanged with the neighbours in backward).
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    a, b = x_range
    Y, Z = sdf_own.shape[1], sdf_own.shape[2]
    if b - a < HALO and world > 1:
        raise ValueError("slab thinner than the halo")
    if extractor is None:
        return _extract_slab_fused(alg, sdf_own, deform_own, a, b, X, isovalue, normalize, group, rank, world)

    sdf_ext = _HaloExchange.apply(sdf_own, rank, world, group) if world > 1 else sdf_own
    def_ext = None
    if deform_own is not None:
        def_ext = _HaloExchange.apply(deform_own, rank, world, group) if world > 1 else deform_own
    lo = a - (HALO if rank > 0 else 0)                      # real x of the first local layer
    verts_l, faces_l, e_pre, f_pre = extractor(sdf_ext, def_ext, isovalue)

    # owned padded layers [A, B) in global padded coords -> local padded layer = xp - lo
    A = a + 1 if rank > 0 else 0
    B = b + 1 if rank < world - 1 else X + 2
    lA, lB = A - lo, B - lo
    k = 3 if alg == "mc" else 4
    e0, e1 = int(e_pre[lA]), int(e_pre[lB])       # crossing edges: MC vertices / DMC quads
    f0, f1 = int(f_pre[lA]), int(f_pre[lB])       # MC triangles / DMC dual vertices
    n_vert = (e1 - e0) if alg == "mc" else (f1 - f0)
    n_face = (f1 - f0) if alg == "mc" else (e1 - e0)
    any_gt = int((sdf_own.detach() > isovalue).any())

    mine = torch.tensor([n_vert, n_face, any_gt], dtype=torch.int64, device=sdf_own.device)
    if world > 1:
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        if dist.get_backend(group) == "gloo" and mine.is_cuda:
            cpu = [g.cpu() for g in gathered]
            dist.all_gather(cpu, mine.cpu(), group=group)
            allc = torch.stack(cpu)
        else:
            dist.all_gather(gathered, mine, group=group)
            allc = torch.stack(gathered).cpu()
    else:
        allc = mine.cpu()[None]
    v_off = int(allc[:rank, 0].sum())
    info = dict(vert_offset=v_off, face_offset=int(allc[:rank, 1].sum()), n_verts_total=int(allc[:, 0].sum()),
                n_faces_total=int(allc[:, 1].sum()), owned_layers=(A, B))
    dev, dt = sdf_own.device, sdf_own.dtype
    # reference early-out on the GLOBAL grid (diso/__init__.py:49): min >= iso <=> no crossing edge anywhere
    if info["n_verts_total"] == 0 or int(allc[:, 2].sum()) == 0:
        info.update(n_verts_total=0, n_faces_total=0, vert_offset=0, face_offset=0)
        return torch.zeros((0, 3), dtype=dt, device=dev), torch.zeros((0, k), dtype=torch.int64, device=dev), info

    # Local ids -> global ids.  Inside the layers a rank owns or references, the local numbering is
    # the global one up to a constant: MC triangles reference vertices of layers [A, B] (layer B is
    # the first layer of the next rank's range), DMC quads reference dual vertices of cell layers
    # [A-1, B-1] (layer A-1 is the last layer of the previous rank's range).
    if alg == "mc":
        verts = verts_l[e0:e1]
        faces = faces_l[f0:f1] - e0 + v_off
    else:
        verts = verts_l[f0:f1]
        faces = faces_l[e0:e1] - f0 + v_off
    shift = torch.zeros(3, dtype=dt, device=dev)
    shift[0] = lo
    verts = verts + shift
    if normalize:
        verts = verts / (torch.tensor([X, Y, Z], dtype=dt, device=dev) - 1)
    return verts, faces, info
