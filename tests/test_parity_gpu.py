"""Parity of the CUDA path (through the public modules, i.e. through the C ABI) against the CPU
oracle on identical seeded inputs.

Bar (BASELINE.json north_star): case index, active-cell set and face connectivity bit-exact;
vertices and gradients within 1e-5 relative (fp32) / 1e-12 (fp64).  We hold vertices to a
stricter bar -- bit-exact -- because the arithmetic order is pinned on both sides; gradients
are compared with a tolerance scaled by the largest gradient magnitude (summation order differs:
the reference uses atomics, the oracle cell order, the kernels a fixed gather order).
"""
import numpy as np
import pytest
import torch

from tests import cases

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
TOL = {torch.float32: 1e-5, torch.float64: 1e-12}


def weights(n, dtype, device="cpu"):
    i = torch.arange(n * 3, dtype=torch.float64).reshape(n, 3)
    return torch.cos(i * 0.6180339887 + 0.25).to(dtype).to(device)


def run_cuda(alg, sdf, deform, iso, normalize, dtype, grad_mode="reference", return_quads=True):
    import diso_b200
    mod = diso_b200.DiffMC(dtype) if alg == "mc" else diso_b200.DiffDMC(dtype, grad_mode=grad_mode)
    s = sdf.to(DEV).requires_grad_(True)
    d = deform.to(DEV).requires_grad_(True) if deform is not None else None
    kw = {} if alg == "mc" else dict(return_quads=return_quads)
    verts, faces = mod(s, d, isovalue=iso, normalize=normalize, **kw)
    out = dict(verts=verts.detach().cpu().numpy(), faces=faces.cpu().numpy(), faces_dtype=faces.dtype)
    if verts.shape[0]:
        (verts * weights(verts.shape[0], dtype, DEV)).sum().backward()
        out["gsdf"] = s.grad.cpu().numpy()
        out["gdef"] = d.grad.cpu().numpy() if d is not None else None
    return out


def close(a, b, dtype, what):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, what
    # non-finite values (a grid dimension of 1 makes `dims - 1` zero, as in the reference) must match exactly
    fin = np.isfinite(b)
    # (inf vs nan is not compared: with a zero divisor both sides are garbage of unspecified kind)
    assert np.array_equal(np.isfinite(a), fin), what + ": non-finite pattern"
    a, b = a[fin], b[fin]
    scale = max(1.0, float(np.abs(b).max()) if b.size else 1.0)
    err = float(np.abs(a - b).max()) if a.size else 0.0
    assert err <= TOL[dtype] * scale, "%s: max err %.3e (scale %.3e)" % (what, err, scale)   # the north-star bar itself (round 1 allowed 4x)


F64 = {"sphere32", "rand_flexi_24", "ragged_5x9x70", "ties_int", "iso_0p37", "tiny_2x2x2"}
PARAMS = [(n, torch.float32) for n in cases.CASES] + [(n, torch.float64) for n in sorted(F64)]


@pytest.mark.parametrize("name,dtype", PARAMS, ids=["%s-%s" % (n, "f32" if d == torch.float32 else "f64") for n, d in PARAMS])
@pytest.mark.parametrize("alg", ["mc", "dmc"])
def test_forward_backward_vs_oracle(oracle, name, dtype, alg):
    sdf, deform, iso = cases.make(name, dtype)
    sn = sdf.numpy()
    dn = deform.numpy() if deform is not None else None
    for normalize in (True, False):
        got = run_cuda(alg, sdf, deform, iso, normalize, dtype)
        ev, ef = oracle.forward(alg, sn, dn, iso, normalize)
        assert got["verts"].shape == ev.shape and got["faces"].shape == ef.shape, (got["verts"].shape, ev.shape, got["faces"].shape, ef.shape)
        if ev.shape[0] == 0:
            assert got["faces_dtype"] == torch.int32  # reference early-out quirk
            continue
        assert got["faces_dtype"] == torch.int64
        assert np.array_equal(got["faces"], ef), "face connectivity differs"
        assert np.array_equal(got["verts"], ev, equal_nan=True), "vertices not bit-identical: max err %.3e" % np.nanmax(np.abs(got["verts"] - ev))
        w = weights(ev.shape[0], dtype).numpy()
        egs, egd = oracle.backward(alg, sn, dn, iso, normalize, w, "reference")
        close(got["gsdf"], egs, dtype, "adj_sdf")
        if dn is not None:
            close(got["gdef"], egd, dtype, "adj_deform")


@pytest.mark.parametrize("name", ["rand_dense_19", "ragged_5x9x70", "ties_int", "boundary_negative"])
def test_dmc_exact_grad_mode_vs_oracle(oracle, name):
    for dtype in (torch.float32, torch.float64):
        sdf, deform, iso = cases.make(name, dtype)
        got = run_cuda("dmc", sdf, deform, iso, True, dtype, grad_mode="exact")
        w = weights(got["verts"].shape[0], dtype).numpy()
        egs, egd = oracle.backward("dmc", sdf.numpy(), deform.numpy(), iso, True, w, "exact")
        close(got["gsdf"], egs, dtype, "adj_sdf")
        close(got["gdef"], egd, dtype, "adj_deform")


@pytest.mark.parametrize("alg", ["mc", "dmc"])
@pytest.mark.parametrize("name", ["sphere64", "rand_dense_19", "ragged_5x9x70", "ties_int", "thin_1x7x33"])
def test_case_index_and_active_cells_bit_exact(oracle, alg, name):
    import diso_b200
    sdf, _, iso = cases.make(name)
    codes = diso_b200.debug_cell_codes(alg, sdf.to(DEV), iso).cpu().numpy().reshape(-1)
    g, _ = oracle.pad_inputs(sdf.numpy(), None, iso)
    r = oracle.raw_forward(alg, g, None, iso)
    used = np.nonzero((codes != 0) & (codes != 255))[0]
    assert np.array_equal(used.astype(np.int32), r["used_index"])
    assert np.array_equal(codes[used], r["used_code"])
    c = diso_b200.extract_counts(alg, sdf.to(DEV), iso)
    assert c["used"] == len(r["used_index"]) and c["verts"] == len(r["verts"]) and c["faces"] == len(r["faces"])


@pytest.mark.parametrize("name", sorted(cases.EMPTY_CASES))
def test_empty_surface_early_out(name):
    import diso_b200
    factory, iso = cases.EMPTY_CASES[name]
    sdf = factory().to(DEV).requires_grad_(True)
    v, f = diso_b200.DiffMC()(sdf, None, iso)
    assert v.shape == (0, 3) and f.shape == (0, 3) and f.dtype == torch.int32 and not v.requires_grad
    for rq in (True, False):
        v, f = diso_b200.DiffDMC()(sdf, None, iso, return_quads=rq)
        assert v.shape == (0, 3) and f.shape == (0, 4) and f.dtype == torch.int32


def test_quad_split_vs_torch_reference_code():
    """diso/__init__.py:118-147 restated verbatim-in-behaviour with torch ops on the GPU is the
    oracle for the split.  The kernel reproduces the summation order of torch's row reductions
    (quad_split.cuh), so the output must be IDENTICAL -- exact ties included (fp32 and fp64)."""
    import torch.nn.functional as F
    import diso_b200
    for dtype in (torch.float32, torch.float64):
        for name in ("rand_dense_19", "roundcube32_def", "sphere32", "ties_int", "rand_dense_40x33x70"):
            sdf, deform, iso = cases.make(name, dtype)
            d = deform.to(DEV) if deform is not None else None
            verts, quads = diso_b200.DiffDMC(dtype)(sdf.to(DEV), d, iso, return_quads=True)
            faces = diso_b200.split_quads(verts, quads)

            def score(cfg):
                out = []
                for tri in cfg:
                    v0, v1, v2 = torch.unbind(verts[quads[:, tri]], dim=-2)
                    c1 = (F.normalize(v1 - v0, dim=-1) * F.normalize(v2 - v0, dim=-1)).sum(-1)
                    c2 = (F.normalize(v2 - v1, dim=-1) * F.normalize(v0 - v1, dim=-1)).sum(-1)
                    c3 = (F.normalize(v0 - v2, dim=-1) * F.normalize(v1 - v2, dim=-1)).sum(-1)
                    out.append(torch.max(torch.stack([c1, c2, c3], -1), -1)[0])
                return torch.max(torch.stack(out, -1), 1)[0]
            a1, a2 = score([[0, 1, 3], [1, 2, 3]]), score([[0, 1, 2], [0, 2, 3]])
            sel = a1 < a2
            ref = torch.cat([quads[sel][:, [0, 1, 3, 1, 2, 3]].view(-1, 3), quads[~sel][:, [0, 1, 2, 0, 2, 3]].view(-1, 3)], 0)
            assert faces.shape == ref.shape and faces.dtype == torch.int64
            assert torch.equal(faces, ref), "%s %s: %d face rows differ from the torch restatement (ties: %d)" % (
                name, dtype, int((faces != ref).any(1).sum()), int((a1 == a2).sum()))


def test_dmc_triangles_default_path(oracle):
    import diso_b200
    for name in ("roundcube32_def", "rand_dense_19", "ties_int"):
        sdf, deform, iso = cases.make(name)
        v, f = diso_b200.DiffDMC()(sdf.to(DEV), deform.to(DEV), iso)  # return_quads=False
        v2, q = diso_b200.DiffDMC()(sdf.to(DEV), deform.to(DEV), iso, return_quads=True)
        assert f.shape == (2 * q.shape[0], 3) and f.dtype == torch.int64 and torch.equal(v, v2)
        ef, _, _ = oracle.split_quads(v.cpu().numpy(), q.cpu().numpy())
        assert np.array_equal(f.cpu().numpy(), ef), "quad split differs from the numpy restatement"


@pytest.mark.parametrize("alg", ["mc", "dmc"])
def test_gradient_of_one_input_only(alg):
    """Only the inputs that require a gradient get one (ctx.needs_input_grad): the backward kernel skips the other
    output entirely (NULL pointer in the C ABI), and the gradient that IS computed equals the one of the full run."""
    import diso_b200
    sdf, deform, iso = cases.make("rand_flexi_24")
    mod = diso_b200.DiffMC() if alg == "mc" else diso_b200.DiffDMC()
    kw = {} if alg == "mc" else dict(return_quads=True)

    def grads(req_s, req_d):
        s = sdf.to(DEV).requires_grad_(req_s)
        d = deform.to(DEV).requires_grad_(req_d)
        v, _ = mod(s, d, iso, **kw)
        (v * weights(v.shape[0], v.dtype, DEV)).sum().backward()
        return s.grad, d.grad
    gs, gd = grads(True, True)
    gs1, gd1 = grads(True, False)
    gs2, gd2 = grads(False, True)
    assert gd1 is None and gs2 is None
    assert torch.equal(gs, gs1) and torch.equal(gd, gd2)
    # retained graph: a second backward over the same forward (the saved state / edge records are only read)
    s = sdf.to(DEV).requires_grad_(True)
    d = deform.to(DEV).requires_grad_(True)
    v, _ = mod(s, d, iso, **kw)
    loss = (v * weights(v.shape[0], v.dtype, DEV)).sum()
    loss.backward(retain_graph=True)
    g1 = s.grad.clone()
    s.grad = None
    loss.backward()
    assert torch.equal(g1, s.grad)
    # forward-only calls keep no edge records
    with torch.no_grad():
        v2, _ = mod(sdf.to(DEV), deform.to(DEV), iso, **kw)
    assert torch.equal(v2, v.detach())


def test_noncontiguous_inputs_and_expanded_grad():
    import diso_b200
    sdf, deform, iso = cases.make("rand_flexi_24")
    s = sdf.to(DEV)
    st = s.permute(2, 1, 0).contiguous().permute(2, 1, 0)  # same values, non-contiguous
    assert not st.is_contiguous()
    m = diso_b200.DiffMC()
    v1, f1 = m(s, None, iso, normalize=False)
    v2, f2 = m(st, None, iso, normalize=False)
    assert torch.equal(v1, v2) and torch.equal(f1, f2)
    s2 = s.clone().requires_grad_(True)
    v, _ = m(s2, None, iso, normalize=False)
    v.sum().backward()  # expanded (non-contiguous) adj_verts: the reference raises here (pybind.cpp:142)
    assert s2.grad is not None and torch.isfinite(s2.grad).all()


def test_two_extractions_interleaved_are_independent():
    """State lives in ctx, not in the module: backward of call A after forward of call B."""
    import diso_b200
    m = diso_b200.DiffMC()
    a = cases.make("sphere32")[0].to(DEV).requires_grad_(True)
    b = cases.make("rand_dense_19")[0].to(DEV).requires_grad_(True)
    va, _ = m(a)
    vb, _ = m(b)
    (va * weights(va.shape[0], torch.float32, DEV)).sum().backward()
    ga = a.grad.clone()
    a.grad = None
    va2, _ = m(a)
    (va2 * weights(va2.shape[0], torch.float32, DEV)).sum().backward()
    assert torch.equal(ga, a.grad)
    del vb
