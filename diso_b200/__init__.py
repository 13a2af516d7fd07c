"""diso_b200 -- B200-native differentiable Marching Cubes / Dual Marching Cubes.

Drop-in for the reference package ``diso`` (``from diso_b200 import DiffMC, DiffDMC``): same
``torch.nn.Module``s, same ``forward`` signatures and return conventions as
/root/reference/diso/__init__.py:9-147, fp32/fp64, autograd into ``grid`` and ``deform``.

Host code is Python/PyTorch (tensor allocation, autograd plumbing, streams); all computation is
in hand-written sm_100a CUDA kernels behind the C ABI of ``include/diso_b200.h``.
"""
import contextlib
import ctypes
import os

import torch
from torch import nn
from torch.autograd import Function

from . import _lib
from ._lib import DisoB200Error  # noqa: F401

__all__ = ["DiffMC", "DiffDMC", "extract_counts", "debug_cell_codes", "split_quads", "layer_prefixes"]
__version__ = "0.1.0"

_DTYPES = {torch.float32: _lib.F32, torch.float64: _lib.F64}


def _ptr(t):
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """cudaStream_t of the current device's current stream.  torch.cuda.current_stream() builds a Stream object through several
    layers of device-index helpers (~8 us a time, three times per forward+backward: a tenth of a 64^3 extraction,
    tools/host_profile.py); the raw getter is one C call."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


_NULL = contextlib.nullcontext()


def _on(device):
    """Context making `device` current -- a no-op object when it already is (torch.cuda.device costs ~5 us a time,
    which is what a 64^3 extraction is made of)."""
    return _NULL if device.index == torch.cuda.current_device() else torch.cuda.device(device)


_state_bytes_cache = {}


def _state_bytes(alg, X, Y, Z):
    key = (alg, X, Y, Z)
    n = _state_bytes_cache.get(key)
    if n is None:
        L = _lib.load()
        n = L.diso_b200_state_bytes(alg, X, Y, Z)
        if n == 0:
            _lib.check(-1 if not L.diso_b200_last_error() else -4)
        if len(_state_bytes_cache) < 4096:
            _state_bytes_cache[key] = n
    return n


def _blocks32(n):
    """Per-edge side arrays (saved edge records, DMC edge adjoints) are stored in groups of 32 edges."""
    return max((n + 31) // 32, 1)


def _check_inputs(grid, deform, dtype):
    # mirrors the checks of the reference glue (src/pybind.cpp:5-11,57-60): CUDA tensors of the
    # extractor's dtype.  Contiguity is not required of the caller (the reference's F.pad makes a
    # contiguous copy, diso/__init__.py:52); we take a contiguous view/copy only when needed.
    if not grid.is_cuda:
        raise DisoB200Error("grid must be a CUDA tensor")
    if grid.dim() != 3:
        raise DisoB200Error("grid must be 3-D [X,Y,Z], got shape %s" % (tuple(grid.shape),))
    if grid.dtype != dtype:
        raise DisoB200Error("grid type must match the extractor dtype (%s), got %s" % (dtype, grid.dtype))
    if deform is not None:
        if not deform.is_cuda or deform.device != grid.device:
            raise DisoB200Error("deform must be a CUDA tensor on the grid's device")
        if deform.dtype != dtype:
            raise DisoB200Error("deformation type must match the extractor dtype (%s), got %s" % (dtype, deform.dtype))
        if tuple(deform.shape) != tuple(grid.shape) + (3,):
            raise DisoB200Error("deform must have shape [X,Y,Z,3], got %s" % (tuple(deform.shape),))


_pinned = {}   # thread id -> pinned int64[COUNT_SLOTS] receiving the count block (one small D2H per forward)


def _pinned_counts():
    import threading
    key = threading.get_ident()
    buf = _pinned.get(key)
    if buf is None:
        t = torch.empty(_lib.COUNT_SLOTS, dtype=torch.int64).pin_memory()
        buf = _pinned[key] = (t, (ctypes.c_int64 * _lib.COUNT_SLOTS).from_address(t.data_ptr()))
    return buf


def _count(alg, grid, isovalue):
    """Phase 1 + the forward's single host sync.  Returns (state tensor, counts list)."""
    L = _lib.load()
    X, Y, Z = grid.shape
    nbytes = _state_bytes(alg, X, Y, Z)
    state = torch.empty(nbytes, dtype=torch.uint8, device=grid.device)
    st = _stream()
    _lib.check(L.diso_b200_count(alg, grid.data_ptr(), _DTYPES[grid.dtype], X, Y, Z, float(isovalue),
                                 state.data_ptr(), nbytes, st))
    pinned, view = _pinned_counts()
    _lib.check(L.diso_b200_read_counts(state.data_ptr(), pinned.data_ptr(), st))   # the one sync
    return state, list(view)


class _Extract(Function):
    """autograd node shared by DiffMC / DiffDMC (reference: DMCFunction / DDMCFunction,
    diso/__init__.py:18-44, 73-98).  Unlike the reference, backward does not re-run the
    forward: the compact rank structure built by phase 1 is saved in ctx as a tensor, which
    keeps batch / activation-checkpoint semantics (nothing lives in the extractor object)."""

    @staticmethod
    def forward(ctx, grid, deform, alg, isovalue, normalize, grad_mode, state, counts, frame=None, aux=None):
        # frame: None, or (x_origin, X_global, id_offset) when `grid` is a slab of a larger grid
        # (diso_b200/parallel.py): vertices come out in the global frame, faces with global ids
        # aux: None, or a dict that receives side outputs; aux["want_quad_flags"] makes the DMC quad kernel decide every
        # quad's diagonal for the triangle split (aux["quad_flags"], one byte per quad) while its four ids are in registers
        L = _lib.load()
        ctx.frame = _lib.Frame(*[int(v) for v in frame]) if frame is not None else None
        X, Y, Z = grid.shape
        k = 3 if alg == _lib.ALG_MC else 4
        n_verts, n_faces = counts[_lib.CNT_VERTS], counts[_lib.CNT_FACES]
        ctx.counts = _lib.counts_array(counts)   # host copy of the count block: sizes the emit / backward launches
        verts = torch.empty((n_verts, 3), dtype=grid.dtype, device=grid.device)
        faces = torch.empty((n_faces, k), dtype=torch.int64, device=grid.device)
        n_edges = n_verts if alg == _lib.ALG_MC else n_faces
        # saved edge records (include/diso_b200.h: edge_rec): only when a gradient can be asked for later
        rec = None
        if ctx.needs_input_grad[0] or (deform is not None and ctx.needs_input_grad[1]):
            ncomp = (5 if deform is not None else 2) + (0 if alg == _lib.ALG_MC else 1)   # {p1 - p0, d0, d1} | {d0, d1}, + quad meta
            rec = torch.empty((_blocks32(n_edges), ncomp, 32), dtype=grid.dtype, device=grid.device)
        args = (grid.data_ptr(), _ptr(deform), _DTYPES[grid.dtype], X, Y, Z, float(isovalue), state.data_ptr(),
                ctypes.cast(ctx.counts, ctypes.c_void_p), int(bool(normalize)), _lib.frame_ptr(ctx.frame))
        if alg == _lib.ALG_MC:
            _lib.check(L.diso_b200_mc_emit(*args, verts.data_ptr(), faces.data_ptr(), _ptr(rec), max(n_edges, 1), _stream()))
        else:
            scratch = torch.empty((max(n_faces, 1), 3), dtype=grid.dtype, device=grid.device)  # edge crossings
            qflags = None
            if aux is not None and aux.get("want_quad_flags"):
                qflags = aux["quad_flags"] = torch.empty(max(n_faces, 1), dtype=torch.uint8, device=grid.device)
            _lib.check(L.diso_b200_dmc_emit(*args, scratch.data_ptr(), verts.data_ptr(), faces.data_ptr(), _ptr(rec),
                                            max(n_edges, 1), _ptr(qflags), _stream()))
        ctx.alg, ctx.isovalue, ctx.normalize, ctx.grad_mode = alg, float(isovalue), bool(normalize), grad_mode
        ctx.n_edges = n_edges
        ctx.has_rec = rec is not None
        if rec is not None:
            # (DMC: the quads themselves are what the backward gathers dL/d dual-vertices with; they are saved as an
            #  output, so modifying them in place before backward raises autograd's usual error)
            if alg == _lib.ALG_MC:
                ctx.save_for_backward(grid, deform, state, rec)
            else:
                ctx.save_for_backward(grid, deform, state, rec, faces)
        else:
            ctx.save_for_backward(grid, deform, state)
        ctx.mark_non_differentiable(faces)
        # do not let autograd allocate + fill a zero "gradient" for the (multi-GB) int64 faces output
        ctx.set_materialize_grads(False)
        return verts, faces

    @staticmethod
    def backward(ctx, adj_verts, adj_faces):
        if ctx.has_rec and ctx.alg != _lib.ALG_MC:
            grid, deform, state, rec, faces = ctx.saved_tensors
        elif ctx.has_rec:
            (grid, deform, state, rec), faces = ctx.saved_tensors, None
        else:
            (grid, deform, state), rec, faces = ctx.saved_tensors, None, None
        L = _lib.load()
        X, Y, Z = grid.shape
        need_grid = ctx.needs_input_grad[0]
        need_deform = deform is not None and ctx.needs_input_grad[1]
        rest = (None,) * 8   # alg, isovalue, normalize, grad_mode, state, counts, frame, aux
        if adj_verts is None:  # verts did not take part in the loss: all gradients are zero
            return (torch.zeros_like(grid) if need_grid else None, torch.zeros_like(deform) if need_deform else None) + rest
        # the reference requires a contiguous adj_verts and raises otherwise (pybind.cpp:142);
        # expanded gradients (e.g. from verts.sum() with normalize=False) are made contiguous here.
        adj_verts = adj_verts.contiguous()
        with _on(grid.device):
            # fully written by the kernel, zeros included; an input that needs no gradient gets no buffer at all
            # (the saved-record backward skips that output; without records both are required by the ABI)
            want_grid = need_grid or rec is None
            want_deform = need_deform or (rec is None and deform is not None)
            adj_grid = torch.empty_like(grid) if want_grid else None
            adj_deform = torch.empty_like(deform) if want_deform else None
            dt = _DTYPES[grid.dtype]
            if ctx.alg == _lib.ALG_MC:
                _lib.check(L.diso_b200_mc_backward(grid.data_ptr(), _ptr(deform), dt, X, Y, Z, ctx.isovalue,
                                                   state.data_ptr(), ctypes.cast(ctx.counts, ctypes.c_void_p),
                                                   adj_verts.data_ptr(), int(ctx.normalize),
                                                   _lib.frame_ptr(ctx.frame), _ptr(rec), max(ctx.n_edges, 1),
                                                   _ptr(adj_grid), _ptr(adj_deform), _stream()))
            else:
                # per-edge adjoints: only the unfused path (no saved records) materialises them
                fused = rec is not None and not os.environ.get("DISO_DMC_BWD_UNFUSED")
                scratch = None if fused else torch.empty((_blocks32(ctx.n_edges), 3, 32), dtype=grid.dtype, device=grid.device)
                _lib.check(L.diso_b200_dmc_backward(grid.data_ptr(), _ptr(deform), dt, X, Y, Z, ctx.isovalue,
                                                    state.data_ptr(), ctypes.cast(ctx.counts, ctypes.c_void_p),
                                                    adj_verts.data_ptr(), int(ctx.normalize), _lib.frame_ptr(ctx.frame),
                                                    ctx.grad_mode, _ptr(rec), max(ctx.n_edges, 1), _ptr(faces) if fused else None, _ptr(scratch),
                                                    _ptr(adj_grid), _ptr(adj_deform), _stream()))
        return (adj_grid if need_grid else None, adj_deform if need_deform else None) + rest


def _run(alg, dtype, grad_mode, grid, deform, isovalue, normalize, want_state=False, slab_mode=False, aux=None):
    _check_inputs(grid, deform, dtype)
    k = 3 if alg == _lib.ALG_MC else 4
    with _on(grid.device):
        g = grid.contiguous()
        d = deform.contiguous() if deform is not None else None
        state, counts = _count(alg, g, isovalue)     # allocates and launches only: nothing autograd could record
        n_verts, n_faces = counts[_lib.CNT_VERTS], counts[_lib.CNT_FACES]
        # diso/__init__.py:49-50,103-104: empty-surface early-out (min >= iso or max <= iso),
        # which returns detached (0,3) verts and INT32 (0,3)/(0,4) faces.
        # (slab_mode: the "max <= iso" half of the test is a property of the GLOBAL grid, decided by the caller)
        if counts[_lib.CNT_EDGES] == 0 or (counts[_lib.CNT_ANY_GT] == 0 and not slab_mode):
            out = (torch.zeros((0, 3), dtype=dtype, device=grid.device),
                   torch.zeros((0, k), dtype=torch.int32, device=grid.device))
            return out + (None,) if want_state else out
        if max(n_verts, n_faces) >= 2 ** 32 - 1:
            raise DisoB200Error("mesh too large for one call (%d verts, %d faces): shard the grid" % (n_verts, n_faces))
        out = _Extract.apply(g, d, alg, float(isovalue), bool(normalize), grad_mode, state, counts, None, aux)
        return out + (state,) if want_state else out


def layer_prefixes(alg, state, shape):
    """Per padded x-layer exclusive prefix sums read from a state tensor (host int64 tensors of
    length X+3): number of crossing edges (== MC vertices == DMC quads) and of MC triangles /
    DMC dual vertices owned by all layers before layer xp.  All output orderings are ascending
    in x, so the items owned by layers [a, b) are the contiguous ranges prefix[a]:prefix[b]."""
    import ctypes
    L = _lib.load()
    alg_id = {"mc": _lib.ALG_MC, "dmc": _lib.ALG_DMC}.get(alg, alg)
    X, Y, Z = shape
    lay = (ctypes.c_int64 * 8)()
    _lib.check(L.diso_b200_state_layout(alg_id, X, Y, Z, lay))
    off_e, off_aux, nch, sx = lay[1], lay[2], lay[5], lay[7]
    idx = torch.arange(0, X + 3, device=state.device, dtype=torch.int64) * sx  # layer starts; last == NCH (totals)
    words = state.view(torch.int32)
    e = words[off_e // 4 + idx * 4].to(torch.int64) & 0xffffffff
    if alg_id == _lib.ALG_MC:
        f = words[off_aux // 4 + idx * 2].to(torch.int64) & 0xffffffff
    else:
        f = words[off_aux // 4 + idx * 4].to(torch.int64) & 0xffffffff
    assert int(idx[-1]) == nch
    return e.cpu(), f.cpu()


def _run_batch(alg, dtype, grad_mode, grids, deforms, isovalue, normalize):
    """B shapes, ONE host sync: all count phases are enqueued first, the B count blocks are read
    back together, then every emit runs (the reference's batch loop, README.md:61-70, pays five
    syncs per shape)."""
    L = _lib.load()
    k = 3 if alg == _lib.ALG_MC else 4
    deforms = list(deforms) if deforms is not None else [None] * len(grids)
    if len(deforms) != len(grids):
        raise DisoB200Error("grids and deforms must have the same length")
    prepared = []
    for g, d in zip(grids, deforms):
        _check_inputs(g, d, dtype)
        if g.device != grids[0].device:
            raise DisoB200Error("all grids of a batch must live on the same device")
        prepared.append((g.contiguous(), d.contiguous() if d is not None else None))
    if not prepared:
        return []
    dev = grids[0].device
    with torch.cuda.device(dev):
        states = []
        with torch.no_grad():
            for g, _ in prepared:
                X, Y, Z = g.shape
                nbytes = L.diso_b200_state_bytes(alg, X, Y, Z)
                state = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                _lib.check(L.diso_b200_count(alg, g.data_ptr(), _DTYPES[g.dtype], X, Y, Z, float(isovalue),
                                             state.data_ptr(), nbytes, _stream()))
                states.append(state)
            heads = torch.stack([s[: 8 * _lib.COUNT_SLOTS].view(torch.int64) for s in states]).cpu().tolist()  # the one sync
        out = []
        for (g, d), state, c in zip(prepared, states, heads):
            if c[_lib.CNT_EDGES] == 0 or c[_lib.CNT_ANY_GT] == 0:
                out.append((torch.zeros((0, 3), dtype=dtype, device=dev), torch.zeros((0, k), dtype=torch.int32, device=dev)))
                continue
            out.append(_Extract.apply(g, d, alg, float(isovalue), bool(normalize), grad_mode, state, c))
        return out


def _grad_mode(name):
    try:
        return {"reference": _lib.GRAD_REFERENCE, "exact": _lib.GRAD_EXACT}[name]
    except KeyError:
        raise ValueError("grad_mode must be 'reference' or 'exact'") from None


class DiffMC(nn.Module):
    """Differentiable Marching Cubes (reference: diso/__init__.py:9-61).

    forward(grid [X,Y,Z], deform [X,Y,Z,3] | None, isovalue=0.0, normalize=True)
        -> verts [V,3] (dtype), faces [F,3] (int64)
    """

    def __init__(self, dtype=torch.float32):
        super().__init__()
        if dtype not in _DTYPES:
            raise DisoB200Error("DiffMC supports torch.float32 / torch.float64, got %s" % (dtype,))
        self.dtype = dtype
        _lib.load()  # fail at construction, not first use, if the CUDA library is missing

    def forward(self, grid, deform=None, isovalue=0.0, normalize=True):
        return _run(_lib.ALG_MC, self.dtype, _lib.GRAD_REFERENCE, grid, deform, isovalue, normalize)

    def forward_batch(self, grids, deforms=None, isovalue=0.0, normalize=True):
        """List of (verts, faces), one per shape, with a single host synchronisation for the batch."""
        return _run_batch(_lib.ALG_MC, self.dtype, _lib.GRAD_REFERENCE, grids, deforms, isovalue, normalize)


class DiffDMC(nn.Module):
    """Differentiable Dual Marching Cubes (reference: diso/__init__.py:64-147).

    forward(grid, deform=None, isovalue=0.0, return_quads=False, normalize=True)
        -> verts [V,3] (dtype), faces [F,3] or quads [Q,4] (int64)

    grad_mode: "reference" (default) reproduces the reference's backward bit-for-bug: in
    cudualmc.cu:957-1005 the dual-vertex cursor is never advanced, so all patches of a cell
    receive the gradient of the cell's first dual vertex.  "exact" is the true adjoint of the
    forward.  They differ only for cells with >= 2 patches.
    """

    def __init__(self, dtype=torch.float32, grad_mode="reference"):
        super().__init__()
        if dtype not in _DTYPES:
            raise DisoB200Error("DiffDMC supports torch.float32 / torch.float64, got %s" % (dtype,))
        self.dtype = dtype
        self.grad_mode = grad_mode
        _grad_mode(grad_mode)
        _lib.load()

    def forward(self, grid, deform=None, isovalue=0.0, return_quads=False, normalize=True):
        aux = None if (return_quads or os.environ.get("DISO_B200_NO_QUAD_FLAGS")) else {"want_quad_flags": True}
        verts, quads = _run(_lib.ALG_DMC, self.dtype, _grad_mode(self.grad_mode), grid, deform, isovalue, normalize, aux=aux)
        if return_quads or quads.shape[0] == 0:
            # (the reference's early-out returns the (0,4) int32 tensor even when triangles were asked for)
            return verts, quads
        return verts, split_quads(verts.detach(), quads, aux.get("quad_flags") if aux else None)

    def forward_batch(self, grids, deforms=None, isovalue=0.0, return_quads=False, normalize=True):
        """List of (verts, faces), one per shape, with a single host synchronisation for the batch."""
        res = _run_batch(_lib.ALG_DMC, self.dtype, _grad_mode(self.grad_mode), grids, deforms, isovalue, normalize)
        if return_quads:
            return res
        return [(v, q if q.shape[0] == 0 else split_quads(v.detach(), q)) for v, q in res]


def split_quads(verts, quads, quad_flags=None):
    """Quad -> triangle split of diso/__init__.py:118-147 (max-min-angle diagonal, config-1 quads
    first) as one fused CUDA pass.  verts [V,3] float32/64, quads [Q,4] int64 -> faces [2Q,3] int64.
    quad_flags: the per-quad diagonal flags the DMC emit already computed (DiffDMC's default path), or None."""
    L = _lib.load()
    verts = verts.contiguous()
    quads = quads.contiguous()
    nq = quads.shape[0]
    faces = torch.empty((2 * nq, 3), dtype=torch.int64, device=quads.device)
    if nq == 0:
        return faces
    with torch.cuda.device(quads.device):
        scratch = torch.empty(L.diso_b200_quad_split_scratch_bytes(nq), dtype=torch.uint8, device=quads.device)
        _lib.check(L.diso_b200_quad_split(verts.data_ptr(), _DTYPES[verts.dtype], quads.data_ptr(), nq, _ptr(quad_flags),
                                          scratch.data_ptr(), faces.data_ptr(), _stream()))
    return faces


def extract_counts(alg, grid, isovalue=0.0):
    """Diagnostics: run phase 1 only; returns dict of counts (verts, faces, edges, used cells)."""
    alg_id = {"mc": _lib.ALG_MC, "dmc": _lib.ALG_DMC}[alg]
    with torch.cuda.device(grid.device):
        _, c = _count(alg_id, grid.contiguous(), isovalue)
    return dict(verts=c[_lib.CNT_VERTS], faces=c[_lib.CNT_FACES], any_gt=c[_lib.CNT_ANY_GT],
                edges=c[_lib.CNT_EDGES], used=c[_lib.CNT_USED])


def debug_cell_codes(alg, grid, isovalue=0.0):
    """Diagnostics / parity tests: dense uint8 case index of every padded cell [X+2,Y+2,Z+2]."""
    L = _lib.load()
    alg_id = {"mc": _lib.ALG_MC, "dmc": _lib.ALG_DMC}[alg]
    X, Y, Z = grid.shape
    with torch.cuda.device(grid.device):
        state, _ = _count(alg_id, grid.contiguous(), isovalue)
        codes = torch.empty((X + 2, Y + 2, Z + 2), dtype=torch.uint8, device=grid.device)
        _lib.check(L.diso_b200_debug_cell_codes(alg_id, X, Y, Z, state.data_ptr(), codes.data_ptr(), _stream()))
    return codes
