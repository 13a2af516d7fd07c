"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle (oracle/libdiso_oracle.so).

Mirrors the *API-level* behaviour of the reference's ``diso/__init__.py`` with numpy:
pad with ``iso+1`` / zero deform (``__init__.py:52-54``), call the padded-frame C restatement,
shift by ``-1`` and normalise by ``dims-1`` (``__init__.py:56-60``), widen faces to int64
(``__init__.py:61``), empty-surface early-out (``__init__.py:49-50``), and for the backward the
chain rule of those two elementwise ops followed by the pad-backward slice.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs
import this module.  The product package (diso_b200) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdiso_oracle.so")
_lib = None


class _Mesh(ctypes.Structure):
    _fields_ = [
        ("n_used", ctypes.c_int64),
        ("n_verts", ctypes.c_int64),
        ("n_faces", ctypes.c_int64),
        ("n_quads", ctypes.c_int64),
        ("verts", ctypes.c_void_p),
        ("faces", ctypes.POINTER(ctypes.c_int32)),
        ("used_index", ctypes.POINTER(ctypes.c_int32)),
        ("used_code", ctypes.POINTER(ctypes.c_uint8)),
        ("scalar_size", ctypes.c_int),
    ]


def build(force=False):
    """Compile the C restatement with the committed Makefile (gcc only)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("diso_oracle.c", "diso_oracle_impl.inc", "diso_tables.h")
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
        for sfx, sc in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            for alg in ("mc", "dmc"):
                f = getattr(L, "oracle_%s_forward_%s" % (alg, sfx))
                f.restype = ctypes.POINTER(_Mesh)
                f.argtypes = [vp, vp, i32, i32, i32, sc]
            f = getattr(L, "oracle_mc_backward_%s" % sfx)
            f.restype = None
            f.argtypes = [vp, vp, i32, i32, i32, sc, vp, vp, vp]
            f = getattr(L, "oracle_dmc_backward_%s" % sfx)
            f.restype = None
            f.argtypes = [vp, vp, i32, i32, i32, sc, vp, vp, vp, i32]
        L.oracle_mesh_free.restype = None
        L.oracle_mesh_free.argtypes = [ctypes.POINTER(_Mesh)]
        _lib = L
    return _lib


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError("oracle supports float32/float64 only")


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def pad_inputs(sdf, deform, iso):
    """diso/__init__.py:52-54: constant pad, value iso+1 (python float, cast to dtype) / 0."""
    dt = sdf.dtype
    g = np.pad(sdf, 1, mode="constant", constant_values=dt.type(float(iso) + 1))
    g = np.ascontiguousarray(g, dtype=dt)
    d = None
    if deform is not None:
        d = np.pad(deform, ((1, 1), (1, 1), (1, 1), (0, 0)), mode="constant", constant_values=0)
        d = np.ascontiguousarray(d, dtype=dt)
    return g, d


def raw_forward(alg, grid_p, deform_p, iso):
    """Padded-frame forward == the reference's L2 ``_C.CUMC*/CUDMC*.forward``.
    Returns dict(verts [V,3] padded frame, faces int32 [F,3|4], used_index, used_code)."""
    L = lib()
    sfx = _sfx(grid_p.dtype)
    DX, DY, DZ = grid_p.shape
    fn = getattr(L, "oracle_%s_forward_%s" % (alg, sfx))
    mp = fn(_ptr(grid_p), _ptr(deform_p), DX, DY, DZ, grid_p.dtype.type(iso))
    m = mp.contents
    k = 3 if alg == "mc" else 4
    V, F, U = m.n_verts, m.n_faces, m.n_used
    ct = ctypes.c_float if sfx == "f32" else ctypes.c_double
    verts = np.ctypeslib.as_array(ctypes.cast(m.verts, ctypes.POINTER(ct)), shape=(max(V, 1) * 3,))[: V * 3]
    verts = verts.reshape(V, 3).copy()
    faces = np.ctypeslib.as_array(m.faces, shape=(max(F, 1) * k,))[: F * k].reshape(F, k).copy()
    used = np.ctypeslib.as_array(m.used_index, shape=(max(U, 1),))[:U].copy()
    code = np.ctypeslib.as_array(m.used_code, shape=(max(U, 1),))[:U].copy()
    L.oracle_mesh_free(mp)
    return dict(verts=verts, faces=faces, used_index=used, used_code=code)


def raw_backward(alg, grid_p, deform_p, iso, adj_verts, grad_mode="reference"):
    """Padded-frame adjoint == the reference's ``_C.*.backward`` into zero-initialised buffers."""
    L = lib()
    sfx = _sfx(grid_p.dtype)
    DX, DY, DZ = grid_p.shape
    adj_verts = np.ascontiguousarray(adj_verts, dtype=grid_p.dtype)
    adj_grid = np.zeros_like(grid_p)
    adj_deform = None if deform_p is None else np.zeros_like(deform_p)
    isoc = grid_p.dtype.type(iso)
    if alg == "mc":
        getattr(L, "oracle_mc_backward_%s" % sfx)(
            _ptr(grid_p), _ptr(deform_p), DX, DY, DZ, isoc, _ptr(adj_verts), _ptr(adj_grid), _ptr(adj_deform))
    else:
        getattr(L, "oracle_dmc_backward_%s" % sfx)(
            _ptr(grid_p), _ptr(deform_p), DX, DY, DZ, isoc, _ptr(adj_verts), _ptr(adj_grid), _ptr(adj_deform),
            0 if grad_mode == "reference" else 1)
    return adj_grid, adj_deform


def _is_empty(sdf, iso):
    # diso/__init__.py:49 / :103
    return bool(sdf.min() >= iso or sdf.max() <= iso)


def forward(alg, sdf, deform=None, isovalue=0.0, normalize=True):
    """API-level forward: (verts [V,3] dtype, faces int64 [F,3] (mc) | quads int64 [Q,4] (dmc)).
    The early-out returns int32 faces like the reference does (__init__.py:50,104)."""
    sdf = np.ascontiguousarray(sdf)
    dt = sdf.dtype
    k = 3 if alg == "mc" else 4
    if _is_empty(sdf, isovalue):
        return np.zeros((0, 3), dt), np.zeros((0, k), np.int32)
    g, d = pad_inputs(sdf, deform, isovalue)
    r = raw_forward(alg, g, d, isovalue)
    verts = r["verts"] - dt.type(1)
    if normalize:
        verts = verts / (np.array(sdf.shape, dtype=dt) - dt.type(1))
    return verts.astype(dt, copy=False), r["faces"].astype(np.int64)


def backward(alg, sdf, deform, isovalue, normalize, adj_verts, grad_mode="reference"):
    """API-level backward: gradients w.r.t. the UNPADDED sdf / deform given dL/dverts (API frame)."""
    sdf = np.ascontiguousarray(sdf)
    dt = sdf.dtype
    g, d = pad_inputs(sdf, deform, isovalue)
    adj = np.ascontiguousarray(adj_verts, dtype=dt)
    if normalize:  # autograd of verts / (dims-1)
        adj = adj / (np.array(sdf.shape, dtype=dt) - dt.type(1))
    ag, ad = raw_backward(alg, g, d, isovalue, adj, grad_mode)
    ag = np.ascontiguousarray(ag[1:-1, 1:-1, 1:-1])
    if ad is not None:
        ad = np.ascontiguousarray(ad[1:-1, 1:-1, 1:-1, :])
    return ag, ad


def split_quads(verts, quads):
    """Restatement of the quad->triangle split of diso/__init__.py:118-147 in numpy (same dtype
    as verts): for both diagonals, max cosine over the 2x3 triangle angles with
    x / max(||x||, 1e-12); config 1 ([0,1,3],[1,2,3]) iff angles1 < angles2; output groups all
    config-1 quads first, then config-2 quads, each in quad order.
    Returns (faces int64 [2Q,3], n_config1, margin [Q] = |angles1 - angles2|); quads with a
    margin near 0 are ties whose diagonal is decided by the last bit of the arithmetic."""
    dt = verts.dtype
    quads = quads.astype(np.int64)

    def sum3(p):
        # torch's CUDA reduction over a row of 3 adds (p0 + p2) + p1, each product rounded on its own
        # (measured on B200, tools/probe_torch_reduce.py); with this order the diagonal choice of the
        # reference is reproduced on every quad, ties included
        return ((p[..., 0] + p[..., 2]).astype(dt) + p[..., 1]).astype(dt)

    def nrm(x):
        n = np.sqrt(sum3((x * x).astype(dt))).astype(dt)
        return (x / np.maximum(n, dt.type(1e-12))[..., None]).astype(dt)

    def tri_max_cos(idx):
        v0, v1, v2 = (verts[quads[:, j]] for j in idx)
        c1 = sum3((nrm(v1 - v0) * nrm(v2 - v0)).astype(dt))
        c2 = sum3((nrm(v2 - v1) * nrm(v0 - v1)).astype(dt))
        c3 = sum3((nrm(v0 - v2) * nrm(v1 - v2)).astype(dt))
        return np.maximum(np.maximum(c1, c2), c3)

    a1 = np.maximum(tri_max_cos([0, 1, 3]), tri_max_cos([1, 2, 3]))
    a2 = np.maximum(tri_max_cos([0, 1, 2]), tri_max_cos([0, 2, 3]))
    sel = a1 < a2
    f1 = quads[sel][:, [0, 1, 3, 1, 2, 3]].reshape(-1, 3)
    f2 = quads[~sel][:, [0, 1, 2, 0, 2, 3]].reshape(-1, 3)
    return np.concatenate([f1, f2], 0), int(sel.sum()), np.abs(a1.astype(np.float64) - a2.astype(np.float64))
