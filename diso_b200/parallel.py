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


_idx_cache = {}


def _prefix_index(alg_id, shape, lA, lB, device):
    """int32-word indices (into the state buffer) of the four per-layer prefix values a rank needs:
    E[lA].base, E[lB].base, aux[lA].base, aux[lB].base (diso_b200_state_layout)."""
    import ctypes
    from diso_b200 import _lib
    key = (alg_id, tuple(shape), lA, lB, str(device))
    idx = _idx_cache.get(key)
    if idx is None:
        L = _lib.load()
        lay = (ctypes.c_int64 * 8)()
        _lib.check(L.diso_b200_state_layout(alg_id, shape[0], shape[1], shape[2], lay))
        off_e, off_aux, sx = lay[1], lay[2], lay[7]
        aw = 2 if alg_id == _lib.ALG_MC else 4      # words per F / P record
        idx = torch.tensor([off_e // 4 + lA * sx * 4, off_e // 4 + lB * sx * 4,
                            off_aux // 4 + lA * sx * aw, off_aux // 4 + lB * sx * aw], dtype=torch.int64, device=device)
        if len(_idx_cache) < 256:
            _idx_cache[key] = idx
    return idx


def _extract_ext(alg, sdf_ext, def_ext, a, b, X, isovalue, normalize, group, rank, world, grad_mode="reference"):
    """The product path: `sdf_ext` / `def_ext` are the slab EXTENDED by its halo layers (already exchanged).
    count -> ONE host read (own counts, the four per-layer prefixes and every rank's totals, gathered on the device)
    -> emit with a frame (include/diso_b200.h: diso_b200_frame), so the kernels write global-frame vertices
    (bit-identical to a single extraction of the whole grid) and global vertex ids directly; what a rank owns are
    contiguous slices (views) of the emitted arrays.  No elementwise pass touches the outputs."""
    import diso_b200
    from diso_b200 import _lib
    alg_id = {"mc": _lib.ALG_MC, "dmc": _lib.ALG_DMC}[alg]
    gm = {"reference": _lib.GRAD_REFERENCE, "exact": _lib.GRAD_EXACT}[grad_mode]
    dev, dt = sdf_ext.device, sdf_ext.dtype
    k = 3 if alg == "mc" else 4
    diso_b200._check_inputs(sdf_ext, def_ext, dt)
    lo = a - (HALO if rank > 0 else 0)                      # global x of the first local layer
    A = a + 1 if rank > 0 else 0                            # owned padded layers [A, B), global padded coords
    B = b + 1 if rank < world - 1 else X + 2
    lA, lB = A - lo, B - lo
    with torch.cuda.device(dev):
        g = sdf_ext.contiguous()
        d = def_ext.contiguous() if def_ext is not None else None
        L = _lib.load()
        with torch.no_grad():
            Xl, Y, Z = g.shape
            nbytes = diso_b200._state_bytes(alg_id, Xl, Y, Z)
            state = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.check(L.diso_b200_count(alg_id, g.data_ptr(), diso_b200._DTYPES[dt], Xl, Y, Z, float(isovalue),
                                         state.data_ptr(), nbytes, diso_b200._stream()))
            head = state[: 8 * _lib.COUNT_SLOTS].view(torch.int64)
            pre = state.view(torch.int32)[_prefix_index(alg_id, (Xl, Y, Z), lA, lB, dev)].to(torch.int64) & 0xffffffff
            e_cnt, f_cnt = pre[1] - pre[0], pre[3] - pre[2]   # crossing edges (MC verts / DMC quads), MC tris / DMC dual verts
            mine_dev = torch.stack(([e_cnt, f_cnt] if alg == "mc" else [f_cnt, e_cnt]) + [head[_lib.CNT_ANY_GT]])
            if world > 1 and dist.get_backend(group) != "gloo":
                gathered = torch.empty((world, 3), dtype=torch.int64, device=dev)
                dist.all_gather_into_tensor(gathered, mine_dev, group=group)
                packed = torch.cat([head, pre, gathered.flatten()]).cpu()          # the one host sync
                allc = packed[_lib.COUNT_SLOTS + 4:].view(world, 3)
            else:
                packed = torch.cat([head, pre]).cpu()
                allc = _gather_counts(mine_dev, group, world)
            counts = packed[: _lib.COUNT_SLOTS].tolist()
            e0, e1, f0, f1 = packed[_lib.COUNT_SLOTS: _lib.COUNT_SLOTS + 4].tolist()
        v0, v1, q0, q1 = (e0, e1, f0, f1) if alg == "mc" else (f0, f1, e0, e1)   # owned vertex / face ranges (local ids)
        # "some value > iso" over the extended slab; the halo layers are other ranks' layers, so the OR over the
        # ranks is the global test of the reference's early-out (diso/__init__.py:49)
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


def _extract_slab_fused(alg, sdf_own, deform_own, a, b, X, isovalue, normalize, group, rank, world, grad_mode="reference"):
    """extract_slab's product path for callers holding only THEIR layers: the halo exchange builds the extended slab
    (one copy of the slab; callers that keep the extended slab themselves avoid it, see SlabField)."""
    sdf_ext = _HaloExchange.apply(sdf_own, rank, world, group) if world > 1 else sdf_own
    def_ext = None
    if deform_own is not None:
        def_ext = _HaloExchange.apply(deform_own, rank, world, group) if world > 1 else deform_own
    return _extract_ext(alg, sdf_ext, def_ext, a, b, X, isovalue, normalize, group, rank, world, grad_mode)


class _HaloGrad(Function):
    """Identity on extended slabs whose halo layers were refreshed in place (one or several tensors of the same rank
    layout, e.g. sdf and deform); backward returns the halo layers' gradients to their owners and adds what the
    neighbours send for OUR boundary layers -- ONE batch of neighbour messages for all tensors.  The incoming
    gradients are updated in place (they are the dense tensors our own backward just produced)."""

    @staticmethod
    def forward(ctx, rank, world, group, *exts):
        ctx.rank, ctx.world, ctx.group = rank, world, group
        out = tuple(e.view_as(e) for e in exts)
        return out if len(out) > 1 else out[0]

    @staticmethod
    def backward(ctx, *gs):
        rank, world, group = ctx.rank, ctx.world, ctx.group
        n_lo = HALO if rank > 0 else 0
        n_hi = HALO if rank < world - 1 else 0
        gs = [None if g is None else (g if g.is_contiguous() else g.contiguous()) for g in gs]
        live = [g for g in gs if g is not None]
        # a tensor whose gradient was not asked for on THIS rank still takes part in the exchange on the others:
        # ranks must agree on the message list, so every rank differentiates the same set of fields (documented)
        spec, recv = [], []
        for g in live:
            n = g.shape[0]
            from_prev = g.new_empty((HALO,) + g.shape[1:]) if n_lo else None
            from_next = g.new_empty((HALO,) + g.shape[1:]) if n_hi else None
            if n_lo:        # my lo halo belongs to prev's last layers; prev's hi halo are my first layers
                spec += [("send", g[:HALO], rank - 1), ("recv", from_prev, rank - 1)]
            if n_hi:
                spec += [("send", g[n - HALO:], rank + 1), ("recv", from_next, rank + 1)]
            recv.append((from_prev, from_next))
        _p2p(spec, group)
        for g, (from_prev, from_next) in zip(live, recv):
            n = g.shape[0]
            if n_lo:
                g[n_lo: n_lo + HALO] += from_prev
                g[:n_lo] = 0
            if n_hi:
                g[n - n_hi - HALO: n - n_hi] += from_next
                g[n - n_hi:] = 0
        return (None, None, None) + tuple(gs)


class SlabField:
    """A rank's slab of one large grid, stored EXTENDED by its halo layers so that no per-step copy of the slab is
    needed: ``ext`` ([n_lo + n + n_hi, Y, Z(, 3)], the tensor to optimise -- make it the leaf) and ``own`` (the view
    of the rank's own layers).  ``refresh()`` re-fills the halo layers from the neighbours in place; gradients that
    land on the halo layers are sent back to their owners inside backward and zeroed locally.
    REQUIREMENT: every rank must run backward through the vertices it got (even an empty set), otherwise the
    neighbours block in the gradient exchange."""

    def __init__(self, own_values, rank, world, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.n_lo = HALO if rank > 0 else 0
        self.n_hi = HALO if rank < world - 1 else 0
        n = own_values.shape[0]
        self.ext = own_values.new_empty((self.n_lo + n + self.n_hi,) + tuple(own_values.shape[1:]))
        self.ext[self.n_lo: self.n_lo + n].copy_(own_values.detach())
        self.n = n

    @property
    def own(self):
        return self.ext[self.n_lo: self.n_lo + self.n]

    def refresh(self):
        if self.world == 1:
            return
        with torch.no_grad():
            _p2p(self._refresh_spec(), self.group)

    def _refresh_spec(self):
        e, n_lo, n = self.ext.detach(), self.n_lo, self.n
        spec = []
        if self.n_lo:
            spec += [("send", e[n_lo: n_lo + HALO], self.rank - 1), ("recv", e[:n_lo], self.rank - 1)]
        if self.n_hi:
            spec += [("send", e[n_lo + n - HALO: n_lo + n], self.rank + 1), ("recv", e[n_lo + n:], self.rank + 1)]
        return spec

    def synced(self):
        """The extended tensor with fresh halos, wired into autograd (use this in the forward)."""
        self.refresh()
        return _HaloGrad.apply(self.rank, self.world, self.group, self.ext) if self.world > 1 else self.ext


def sync_fields(fields, refresh=True):
    """Several SlabFields of one rank layout (e.g. sdf and deform) at once: ONE batch of neighbour messages refreshes all
    their halos, and one joint autograd node returns all their halo gradients in one batch during backward.
    refresh=False: the halos are known to be current (a second extraction from unchanged fields in the same step)."""
    fields = [f for f in fields if f is not None]
    f0 = fields[0]
    if f0.world == 1:
        return [f.ext for f in fields]
    if refresh:
        with torch.no_grad():
            spec = []
            for f in fields:
                spec += f._refresh_spec()
            _p2p(spec, f0.group)
    out = _HaloGrad.apply(f0.rank, f0.world, f0.group, *[f.ext for f in fields])
    return list(out) if isinstance(out, tuple) else [out]


def extract_slab_ext(alg, sdf_field, deform_field, x_range, X, isovalue=0.0, normalize=True, group=None, grad_mode="reference",
                     refresh=True):
    """extract_slab for :class:`SlabField` inputs (the extended slab is the leaf: no per-step copy of the slab,
    one host synchronisation per extraction, one batch of halo messages per direction).  Same return values as
    :func:`extract_slab`.  refresh=False skips the halo refresh (fields unchanged since the last refresh)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    a, b = x_range
    exts = sync_fields([sdf_field, deform_field], refresh)
    sdf_ext = exts[0]
    def_ext = exts[1] if deform_field is not None else None
    return _extract_ext(alg, sdf_ext, def_ext, a, b, X, isovalue, normalize, group, rank, world, grad_mode)


def extract_slab(alg, sdf_own, deform_own, x_range, X, isovalue=0.0, normalize=True, group=None, extractor=None):
    """Slab-sharded extraction.  Every rank of `group` calls this with ITS slab
    ``sdf_own = sdf[a:b]`` (and ``deform_own = deform[a:b]`` or None), ``x_range = (a, b)`` from
    :func:`plan_slabs` and the global ``X``.

    Returns ``(verts, faces, info)``: the vertices this rank owns ([n,3], global frame, divided by
    the global ``dims-1`` if ``normalize``), the faces it owns with GLOBAL vertex ids (MC [F,3]
    triangles, DMC [Q,4] quads; int64) and ``info`` with the global offsets / totals.  The global
    mesh is the rank-order concatenation.  Gradients flow to ``sdf_own`` / ``deform_own`` of every
    rank (halo contributions are exchanged with the neighbours in backward).

    REQUIREMENT: the halo exchange is a collective in BOTH directions.  Every rank must run backward through the
    ``verts`` it got -- also a rank that owns nothing (its empty ``verts`` stays attached to the exchange for exactly
    this reason) -- otherwise its neighbours block in the gradient exchange.
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
