"""The emit / adjoint kernels exist in two tile flavours (diso_b200/csrc/compact.cuh): "dense" tiles
of 64 consecutive chunks and "listed" tiles over the ordered active-chunk list (sparse surfaces).
The host picks one from the counts; both must produce identical bits.  Called through the C ABI:
counts_host == NULL forces the dense flavour, the counts read back after phase 1 select the listed
one whenever fewer than 75 % of the chunks are active.  The same switch selects the backward's sparse path
(zero fill + list of touched blocks + persistent grid) against its dense one-CTA-per-block launch."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def blobs(n, seed):
    """A few small spheres scattered in an otherwise empty grid: scattered (non-contiguous) tiles."""
    g = torch.Generator().manual_seed(seed)
    ax = torch.arange(n, dtype=torch.float32)
    x, y, z = torch.meshgrid(ax, ax, ax, indexing="ij")
    sdf = torch.full((n, n, n), 10.0)
    for _ in range(7):
        c = torch.rand(3, generator=g) * (n - 8) + 4
        r = 2.0 + 3.0 * float(torch.rand(1, generator=g))
        sdf = torch.minimum(sdf, ((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2).sqrt() - r)
    return sdf + 0.013 * torch.rand(sdf.shape, generator=g)


@pytest.mark.parametrize("alg", ["mc", "dmc"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("use_def", [True, False])
def test_dense_and_listed_tiles_agree(alg, dtype, use_def):
    import diso_b200
    from diso_b200 import _lib
    L = _lib.load()
    n = 70
    sdf = blobs(n, 3).to(dtype).to(DEV)
    deform = (0.4 * torch.tanh(torch.randn(n, n, n, 3, generator=torch.Generator().manual_seed(5)))).to(dtype).to(DEV) if use_def else None
    dptr = lambda t: None if t is None else t.data_ptr()
    alg_id = _lib.ALG_MC if alg == "mc" else _lib.ALG_DMC
    dt = _lib.F32 if dtype == torch.float32 else _lib.F64
    state, counts = diso_b200._count(alg_id, sdf, 0.0)
    lay = (ctypes.c_int64 * 8)()
    _lib.check(L.diso_b200_state_layout(alg_id, n, n, n, lay))
    nch = lay[5]
    assert 0 < counts[_lib.CNT_EDGE_CHUNKS] * 4 < nch * 3 and 0 < counts[_lib.CNT_CELL_CHUNKS] * 4 < nch * 3, "input must select the listed flavour"
    assert counts[_lib.CNT_EDGE_CHUNKS] * 8 < nch, "input must select the sparse (zero fill + touched-block list) backward"
    nv, nf = counts[_lib.CNT_VERTS], counts[_lib.CNT_FACES]
    k = 3 if alg == "mc" else 4
    ne = nv if alg == "mc" else nf
    st = torch.cuda.current_stream().cuda_stream
    per_path = {}
    for use_rec in (False, True):      # v1 backward (re-gathers sdf / deform)  |  v2 backward from saved edge records
        res = []
        for ch in (ctypes.cast(_lib.counts_array(counts), ctypes.c_void_p), None):
            verts = torch.full((nv, 3), float("nan"), dtype=dtype, device=DEV)
            faces = torch.full((nf, k), -1, dtype=torch.int64, device=DEV)
            adj_s = torch.full_like(sdf, float("nan"))
            adj_d = torch.full_like(deform, float("nan")) if use_def else None
            ncomp = (5 if use_def else 2) + (0 if alg == "mc" else 1)      # include/diso_b200.h: edge_rec
            rec = torch.full(((ne + 31) // 32, ncomp, 32), float("nan"), dtype=dtype, device=DEV) if use_rec else None
            rp = rec.data_ptr() if use_rec else None
            w = torch.cos(torch.arange(nv * 3, dtype=torch.float64).reshape(nv, 3) * 0.618).to(dtype).to(DEV)
            common = (sdf.data_ptr(), dptr(deform), dt, n, n, n, 0.0, state.data_ptr())
            keep = lambda: [adj_s.clone()] + ([adj_d.clone()] if use_def else [])
            grads = []
            if alg == "mc":
                _lib.check(L.diso_b200_mc_emit(*common, ch, 1, None, verts.data_ptr(), faces.data_ptr(), rp, ne, st))
                _lib.check(L.diso_b200_mc_backward(*common, ch, w.data_ptr(), 1, None, rp, ne, adj_s.data_ptr(), dptr(adj_d), st))
                grads += keep()
            else:
                scratch = torch.empty(((ne + 31) // 32 * 32, 3), dtype=dtype, device=DEV)
                _lib.check(L.diso_b200_dmc_emit(*common, ch, 1, None, scratch.data_ptr(), verts.data_ptr(), faces.data_ptr(), rp, ne, None, st))
                for gm in (_lib.GRAD_REFERENCE, _lib.GRAD_EXACT):
                    _lib.check(L.diso_b200_dmc_backward(*common, ch, w.data_ptr(), 1, None, gm, rp, ne, faces.data_ptr() if use_rec else None, scratch.data_ptr(), adj_s.data_ptr(), dptr(adj_d), st))
                    grads += keep()
            torch.cuda.synchronize()
            if use_rec:
                flat = rec.permute(1, 0, 2).reshape(ncomp, -1)[:(5 if use_def else 2), :ne]
                assert bool(torch.isfinite(flat).all()), "edge records not fully written"
            res.append([t.cpu().numpy() for t in [verts, faces] + grads])
        names = ("adj_sdf", "adj_deform", "adj_sdf(exact)", "adj_deform(exact)") if use_def else ("adj_sdf", "adj_sdf(exact)")
        for a, b, what in zip(res[0], res[1], ("verts", "faces") + names):
            assert not np.isnan(a.astype(np.float64)).any() and (a != -1).any(), what + ": not fully written"
            assert np.array_equal(a, b), what + ": listed and dense tile flavours differ (records: %s)" % use_rec
        per_path[use_rec] = res[0]
    # the two backward designs agree to rounding (different, but each fixed, summation order)
    tol = 1e-5 if dtype == torch.float32 else 1e-13
    for a, b in zip(per_path[False][:2], per_path[True][:2]):
        assert np.array_equal(a, b)
    for a, b in zip(per_path[False][2:], per_path[True][2:]):
        assert np.abs(a.astype(np.float64) - b).max() <= tol * max(1.0, np.abs(b).max())


def test_sparse_and_dense_backward_agree_at_scale():
    """Sphere 448^3 (a smooth surface: ~10 % of the chunks own a crossing edge): the zero-fill + touched-block
    path and the one-CTA-per-block path must write identical gradients, zeros included."""
    import diso_b200
    from diso_b200 import _lib, synthetic as syn
    L = _lib.load()
    n = 448
    sdf = syn.sphere_sdf(n).to(DEV)
    deform = syn.random_deform(n, 3).to(DEV)
    state, counts = diso_b200._count(_lib.ALG_MC, sdf, 0.0)
    lay = (ctypes.c_int64 * 8)()
    _lib.check(L.diso_b200_state_layout(_lib.ALG_MC, n, n, n, lay))
    assert counts[_lib.CNT_EDGE_CHUNKS] * 8 < lay[5]
    nv = counts[_lib.CNT_VERTS]
    w = torch.cos(torch.arange(nv * 3, dtype=torch.float64, device=DEV).reshape(nv, 3) * 0.618).float()
    st = torch.cuda.current_stream().cuda_stream
    verts = torch.empty((nv, 3), device=DEV)
    faces = torch.empty((counts[_lib.CNT_FACES], 3), dtype=torch.int64, device=DEV)
    rec = torch.empty(((nv + 31) // 32, 5, 32), device=DEV)
    chc = ctypes.cast(_lib.counts_array(counts), ctypes.c_void_p)
    _lib.check(L.diso_b200_mc_emit(sdf.data_ptr(), deform.data_ptr(), _lib.F32, n, n, n, 0.0, state.data_ptr(), chc, 1, None,
                                   verts.data_ptr(), faces.data_ptr(), rec.data_ptr(), nv, st))
    ref_out = None
    for rp in (None, rec.data_ptr()):       # v1 | v2 (saved records)
        outs = []
        for ch in (chc, None):
            adj_s = torch.full_like(sdf, float("nan"))
            adj_d = torch.full_like(deform, float("nan"))
            _lib.check(L.diso_b200_mc_backward(sdf.data_ptr(), deform.data_ptr(), _lib.F32, n, n, n, 0.0, state.data_ptr(), ch,
                                               w.data_ptr(), 1, None, rp, nv, adj_s.data_ptr(), adj_d.data_ptr(), st))
            torch.cuda.synchronize()
            outs.append((adj_s, adj_d))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
        assert bool(torch.isfinite(outs[0][1]).all()) and float(outs[0][1].abs().sum()) > 0
        assert int((outs[0][0] != 0).sum()) < sdf.numel() // 10   # mostly zeros, all written
        if ref_out is None:
            ref_out = outs[0]
        else:
            for a, b in zip(ref_out, outs[0]):
                assert float((a - b).abs().max()) <= 1e-5 * max(1.0, float(b.abs().max()))
