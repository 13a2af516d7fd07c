"""Compatibility shim for the reference's low-level extension module ``diso._C``
(/root/reference/src/pybind.cpp:419-447): classes ``CUMCFloat``, ``CUMCDouble``, ``CUDMCFloat``,
``CUDMCDouble`` with the reference's calling convention

    verts, faces_int32 = obj.forward(grid_padded[, deform_padded], iso)
    obj.backward(grid_padded[, deform_padded], iso, adj_verts, adj_grid[, adj_deform])   # accumulates in place

for code that drives the extractor below the ``DiffMC`` / ``DiffDMC`` modules.  Semantics kept:
inputs are the caller's ALREADY PADDED tensors (the reference applies no pad at this level and its
kernels rely on the iso+1 shell), vertices come back in the frame of that
tensor with no shift / normalisation, faces are int32, ``backward`` uses the state of the LAST
``forward`` on the object (pybind.cpp:16-39) and adds into caller-zeroed adjoints.

Implementation: the C ABI works on unpadded grids with a virtual iso+1 shell.  Feeding it the
caller's padded tensor adds one more (inert) shell -- no crossing edge can touch it because the
caller's own shell is already >= iso -- so the mesh is the reference's; coordinates are formed one
lattice unit further from the origin and shifted back, hence equal to the reference's to 1 ulp
rather than bit-for-bit.  Input checks mirror pybind.cpp:5-11,57-60 (CUDA, contiguous, dtype).

Limit: that equivalence needs the caller's boundary layer to be >= iso (what diso/__init__.py:52 guarantees for
every call that goes through DiffMC / DiffDMC).  A raw grid whose boundary dips below iso makes the reference read
out of bounds / leave the surface open, while the virtual shell here would close it with extra faces -- different
meshes.  ``forward`` therefore checks the six boundary faces and raises instead of returning a different mesh
(``DISO_B200_C_NO_BOUNDARY_CHECK=1`` skips the check and its small host synchronisation).
"""
import os

import ctypes

import torch

from . import _lib
from . import _count, _Extract, DisoB200Error


class _Base:
    _alg = None
    _dtype = None

    def __init__(self):
        _lib.load()
        self._last = None  # (grid, deform, iso, verts graph) of the last forward

    def _check(self, name, t):
        if not t.is_cuda:
            raise DisoB200Error("%s must be a CUDA tensor" % name)
        if not t.is_contiguous():
            raise DisoB200Error("%s must be contiguous" % name)
        if t.dtype != self._dtype:
            raise DisoB200Error("%s type must match the %s class" % (name, "mc" if self._alg == _lib.ALG_MC else "dmc"))

    def forward(self, grid, *args):
        deform, iso = (None, args[0]) if len(args) == 1 else (args[0], args[1])
        self._check("grid", grid)
        if deform is not None:
            self._check("deform", deform)
        k = 3 if self._alg == _lib.ALG_MC else 4
        if not os.environ.get("DISO_B200_C_NO_BOUNDARY_CHECK") and grid.numel():
            lo = min(float(f.min()) for f in (grid[0], grid[-1], grid[:, 0], grid[:, -1], grid[:, :, 0], grid[:, :, -1]))
            if lo < float(iso):
                raise DisoB200Error("diso_b200._C expects the reference's padded input (boundary layer >= iso, diso/__init__.py:52); "
                                    "this grid's boundary goes down to %g < iso = %g: pad it, or call DiffMC / DiffDMC" % (lo, float(iso)))
        with torch.cuda.device(grid.device), torch.no_grad():
            state, counts = _count(self._alg, grid, iso)
            nv, nf = counts[_lib.CNT_VERTS], counts[_lib.CNT_FACES]
            self._last = (state, nv, nf, _lib.counts_array(counts))
            if counts[_lib.CNT_EDGES] == 0:
                return (torch.zeros((0, 3), dtype=self._dtype, device=grid.device),
                        torch.zeros((0, k), dtype=torch.int32, device=grid.device))
            verts, faces = _Extract.apply(grid, deform, self._alg, float(iso), False, _lib.GRAD_REFERENCE, state, counts)
        # normalize=False output is (padded-frame position - 1) of OUR frame == the caller's frame
        return verts, faces.to(torch.int32)

    def backward(self, grid, *args):
        if len(args) == 3:
            deform, (iso, adj_verts, adj_grid), adj_deform = None, args, None
        else:
            deform, iso, adj_verts, adj_grid, adj_deform = args
        for name, t in (("adj_verts", adj_verts), ("adj_grid", adj_grid)) + ((("adj_deform", adj_deform),) if adj_deform is not None else ()):
            self._check(name, t)
        if self._last is None:
            raise DisoB200Error("backward called before forward")
        state, nv, nf, counts_c = self._last
        if nv == 0 or (self._alg == _lib.ALG_DMC and nf == 0):
            return
        L = _lib.load()
        X, Y, Z = grid.shape
        dt = _lib.F32 if self._dtype == torch.float32 else _lib.F64
        st = torch.cuda.current_stream().cuda_stream
        with torch.cuda.device(grid.device):
            g_grid = torch.empty_like(grid)
            g_def = torch.empty_like(deform) if deform is not None else None
            p = lambda t: None if t is None else t.data_ptr()
            if self._alg == _lib.ALG_MC:
                _lib.check(L.diso_b200_mc_backward(grid.data_ptr(), p(deform), dt, X, Y, Z, float(iso), state.data_ptr(),
                                                   ctypes.cast(counts_c, ctypes.c_void_p),
                                                   adj_verts.data_ptr(), 0, None, None, 0, g_grid.data_ptr(), p(g_def), st))
            else:
                scratch = torch.empty((max(nf, 1), 3), dtype=self._dtype, device=grid.device)
                _lib.check(L.diso_b200_dmc_backward(grid.data_ptr(), p(deform), dt, X, Y, Z, float(iso), state.data_ptr(),
                                                    ctypes.cast(counts_c, ctypes.c_void_p),
                                                    adj_verts.data_ptr(), 0, None, _lib.GRAD_REFERENCE, None, 0, None, scratch.data_ptr(),
                                                    g_grid.data_ptr(), p(g_def), st))
            adj_grid.add_(g_grid)          # the reference accumulates into the caller's buffers (atomicAdd)
            if adj_deform is not None:
                adj_deform.add_(g_def)


class CUMCFloat(_Base):
    _alg, _dtype = _lib.ALG_MC, torch.float32


class CUMCDouble(_Base):
    _alg, _dtype = _lib.ALG_MC, torch.float64


class CUDMCFloat(_Base):
    _alg, _dtype = _lib.ALG_DMC, torch.float32


class CUDMCDouble(_Base):
    _alg, _dtype = _lib.ALG_DMC, torch.float64
