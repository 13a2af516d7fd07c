"""Side-by-side parity with the UNMODIFIED reference CUDA build at the BASELINE sizes (C3 256^3, C4 512^3 with
deform in fp32 AND fp64), plus one >= 1024^3 single call checked against independent sub-grid extractions.

Reference path compared against: /root/reference/diso/__init__.py:48-61,102-147 -> src/cumc.cu:651-732,
src/cudualmc.cu:1058-1128 (built unmodified into baseline/_ref).  Bar: faces `torch.equal`, vertices bit-identical,
gradients within the north-star tolerance of the GRADIENT SCALE; the observed errors (ours and the reference's own,
both against an fp64 run) are printed and appended to gpurun_out/parity_large.jsonl so the bound is set from data.
"""
import json
import os

import pytest
import torch

from diso_b200 import synthetic as syn
from tests.refload import load_reference

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# north star: 1e-5 relative (fp32), 1e-12 (fp64), relative to the gradient scale (gradients are sums of up to 6 (MC) /
# 24 (DMC) terms whose order differs between implementations: the reference accumulates with atomics).
# Set from data (round 2, B200, profiles/r2_parity_large.md): ours vs the reference <= 2.5e-7 (fp32) / 5e-16 (fp64) on
# every BASELINE config, while BOTH differ from an fp64 run by 6e-6 .. 1.3e-5 -- that part is the fp32 rounding of
# `coordinate + deform` at coordinates of a few hundred, which both implementations commit identically.
GRAD_TOL = {torch.float32: 1e-6, torch.float64: 1e-14}


@pytest.fixture(scope="module")
def ref():
    m = load_reference()
    if m is None:
        pytest.skip("reference build baseline/_ref not available")
    return m


def _log(rec):
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_large.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass
    print(json.dumps(rec))


def _weights(n, dtype):
    g = torch.Generator(device=DEV).manual_seed(11)
    return torch.rand((n, 3), generator=g, device=DEV, dtype=torch.float64).to(dtype)


def _run(mod, sdf, deform, w=None, **kw):
    s = sdf.clone().requires_grad_(True)
    d = deform.clone().requires_grad_(True) if deform is not None else None
    v, f = mod(s, d, **kw)
    if w is None:
        w = _weights(v.shape[0], v.dtype)
    (v * w).sum().backward()
    return v.detach(), f, s.grad, (d.grad if d is not None else None), w


def _relerr(a, b):
    scale = max(1.0, float(b.abs().max()))
    return float((a.double() - b.double()).abs().max()) / scale


CASES = [
    # (size, kind, use deform, dtypes)
    (256, "flexi", True, ("f32",)),
    (256, "sparse", False, ("f32",)),
    (256, "dense", True, ("f32",)),
    (512, "flexi", True, ("f32", "f64")),
]
PARAMS = [(n, kind, dfm, dt, alg) for (n, kind, dfm, dts) in CASES for dt in dts for alg in ("mc", "dmc")]


@pytest.mark.parametrize("n,kind,use_def,dt,alg", PARAMS, ids=["%d-%s-%s-%s" % (p[0], p[1], p[3], p[4]) for p in PARAMS])
def test_side_by_side_headline(ref, n, kind, use_def, dt, alg):
    import diso_b200
    dtype = torch.float32 if dt == "f32" else torch.float64
    sdf = syn.random_sdf(n, kind, 0, dtype).to(DEV)
    deform = syn.random_deform(n, 1, dtype).to(DEV) if use_def else None
    kw = {} if alg == "mc" else dict(return_quads=True)
    ours = diso_b200.DiffMC(dtype) if alg == "mc" else diso_b200.DiffDMC(dtype)
    theirs = ref.DiffMC(dtype) if alg == "mc" else ref.DiffDMC(dtype)
    va, fa, gsa, gda, w = _run(ours, sdf, deform, **kw)
    vb, fb, gsb, gdb, _ = _run(theirs, sdf, deform, w=w, **kw)
    assert fa.dtype == fb.dtype and fa.shape == fb.shape
    assert torch.equal(fa, fb), "face connectivity differs from the reference"
    nbad = int((va != vb).any(1).sum())
    assert nbad == 0, "%d vertices differ from the reference in the last bits (max %.3e)" % (nbad, float((va - vb).abs().max()))
    rec = dict(test="side_by_side", n=n, kind=kind, dtype=dt, alg=alg, deform=use_def, verts=int(va.shape[0]), faces=int(fa.shape[0]),
               gsdf_ours_vs_ref=_relerr(gsa, gsb), gdef_ours_vs_ref=_relerr(gda, gdb) if use_def else None)
    del va, vb, fa, fb
    if dt == "f32":
        # ground truth: the fp64 path on the same (upcast) inputs; shows how much of |ours - reference| is the
        # reference's own rounding / atomics-order noise
        hi = diso_b200.DiffMC(torch.float64) if alg == "mc" else diso_b200.DiffDMC(torch.float64)
        _, _, gst, gdt, _ = _run(hi, sdf.double(), deform.double() if use_def else None, w=w.double(), **kw)
        rec.update(gsdf_ours_vs_f64=_relerr(gsa, gst), gsdf_ref_vs_f64=_relerr(gsb, gst))
        if use_def:
            rec.update(gdef_ours_vs_f64=_relerr(gda, gdt), gdef_ref_vs_f64=_relerr(gdb, gdt))
    _log(rec)
    tol = GRAD_TOL[dtype]
    assert rec["gsdf_ours_vs_ref"] <= tol, rec
    if use_def:
        assert rec["gdef_ours_vs_ref"] <= tol, rec
    if dt == "f32":
        # against the fp64 ground truth we must not be worse than the reference itself (+ the ours-vs-reference bar)
        assert rec["gsdf_ours_vs_f64"] <= rec["gsdf_ref_vs_f64"] + tol, rec
        if use_def:
            assert rec["gdef_ours_vs_f64"] <= rec["gdef_ref_vs_f64"] + tol, rec


@pytest.mark.parametrize("n,kind", [(256, "flexi"), (512, "flexi")])
def test_default_dmc_triangles_headline(ref, n, kind):
    """DiffDMC's default path (return_quads=False, /root/reference/diso/__init__.py:117-147) at the BASELINE sizes:
    the triangle list must equal the reference's; quads whose two diagonals tie to the last bit may flip and are counted."""
    import diso_b200
    sdf = syn.random_sdf(n, kind, 0).to(DEV)
    deform = syn.random_deform(n, 1).to(DEV)
    with torch.no_grad():
        va, fa = diso_b200.DiffDMC()(sdf, deform)
        vb, fb = ref.DiffDMC()(sdf, deform)
    assert torch.equal(va, vb) and fa.shape == fb.shape and fa.dtype == fb.dtype
    same = torch.equal(fa, fb)
    flipped = 0
    if not same:
        # config-1 quads come first in both lists; the number of rows that differ bounds the flipped quads
        with torch.no_grad():
            _, q = diso_b200.DiffDMC()(sdf, deform, return_quads=True)
        V = va.shape[0]

        def keyset(f):
            return torch.sort(f[:, 0] * V * V + f[:, 1] * V + f[:, 2])[0]
        want = q[:, 0] * V * V + q[:, 1] * V + q[:, 3]     # [q0,q1,q3] present <=> config 1

        def choice(f):
            have = keyset(f)
            pos = torch.searchsorted(have, want).clamp(max=have.numel() - 1)
            return have[pos] == want
        flipped = int((choice(fa) != choice(fb)).sum())
    _log(dict(test="dmc_default_triangles", n=n, kind=kind, quads=int(fa.shape[0] // 2), identical=same, flipped_diagonals=flipped))
    assert flipped == 0, "%d quads pick the other diagonal than the reference" % flipped
    assert same


def _sub_extract(alg_id, sdf_sub, x0, X, id_offset):
    """Independent extraction of the x-range [x0, x0 + n) of a larger grid in the GLOBAL frame (frame = the C ABI's
    diso_b200_frame): returns verts, faces, per-layer prefixes of the sub-grid."""
    import diso_b200
    from diso_b200 import _lib
    with torch.no_grad():
        state, counts = diso_b200._count(alg_id, sdf_sub, 0.0)
        e_pre, f_pre = diso_b200.layer_prefixes(alg_id, state, tuple(sdf_sub.shape))
        nv = counts[_lib.CNT_VERTS]
        verts, faces = diso_b200._Extract.apply(sdf_sub, None, alg_id, 0.0, True, _lib.GRAD_REFERENCE, state, counts,
                                                (x0, X, id_offset - nv))   # ids such that the LAST local vertex gets id_offset - 1
    return verts, faces, e_pre, f_pre


@pytest.mark.parametrize("alg", ["mc", "dmc"])
def test_1024_single_call(alg):
    """One 1024^3 call (the reference overflows its int32 scans here, /root/reference/src/cumc.cu:717-723):
    580 M vertices, 1.03 G triangles -> element offsets beyond 2^31.  The head and the tail of the outputs must be
    bit-identical to independent extractions of the first / last x-layers; ids must cover [0, V)."""
    import diso_b200
    from diso_b200 import _lib
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 90e9:
        pytest.skip("needs ~70 GB of device memory")
    N = 1024
    alg_id = _lib.ALG_MC if alg == "mc" else _lib.ALG_DMC
    g = torch.Generator(device=DEV).manual_seed(5)
    sdf = torch.rand((N, N, N), generator=g, device=DEV, dtype=torch.float32) - 0.1
    with torch.no_grad():
        verts, faces, state = diso_b200._run(alg_id, torch.float32, _lib.GRAD_REFERENCE, sdf, None, 0.0, True, want_state=True)
    V, Fn = verts.shape[0], faces.shape[0]
    k = faces.shape[1]
    assert faces.numel() > 2 ** 31, "the test is meant to cross the 2^31-element line"
    # independent count of the crossing edges from the sign field (== MC vertices == DMC quads)
    b = torch.nn.functional.pad(sdf >= 0, (1, 1, 1, 1, 1, 1), value=True)
    n_edges = int((b[1:] != b[:-1]).sum()) + int((b[:, 1:] != b[:, :-1]).sum()) + int((b[:, :, 1:] != b[:, :, :-1]).sum())
    del b
    assert (V if alg == "mc" else Fn) == n_edges
    e_pre, f_pre = diso_b200.layer_prefixes(alg_id, state, (N, N, N))
    v_pre, q_pre = (e_pre, f_pre) if alg == "mc" else (f_pre, e_pre)     # per padded layer: vertex / face prefix
    assert int(v_pre[-1]) == V and int(q_pre[-1]) == Fn
    # id range + the last rows are written
    lo, hi = int(faces.min()), int(faces.max())
    assert lo == 0 and hi == V - 1
    assert int(faces[-1].max()) < V and bool(torch.isfinite(verts[-1]).all())
    # ---- tail: the last D layers, extracted on their own (halo of 3 layers at the cut) ----------------------------
    D, H = 40, 3
    x0 = N - D - H
    sv, sf, se, sfp = _sub_extract(alg_id, sdf[x0:].contiguous(), x0, N, V)
    sv_pre, sq_pre = (se, sfp) if alg == "mc" else (sfp, se)
    lA = H + 1                                    # local padded layer of global layer x0 + H (padded: x0 + H + 1)
    gA = x0 + H + 1
    nv_tail, nq_tail = V - int(v_pre[gA]), Fn - int(q_pre[gA])
    assert sv.shape[0] - int(sv_pre[lA]) == nv_tail and sf.shape[0] - int(sq_pre[lA]) == nq_tail
    assert torch.equal(verts[V - nv_tail:], sv[sv.shape[0] - nv_tail:]), "tail vertices differ"
    assert torch.equal(faces[Fn - nq_tail:], sf[sf.shape[0] - nq_tail:]), "tail faces differ"
    del sv, sf
    # ---- head: the first D layers --------------------------------------------------------------------------------
    n_sub = D + H
    sub = sdf[:n_sub].contiguous()
    with torch.no_grad():
        st2, c2 = diso_b200._count(alg_id, sub, 0.0)
        he, hf = diso_b200.layer_prefixes(alg_id, st2, tuple(sub.shape))
        hv, hfaces = diso_b200._Extract.apply(sub, None, alg_id, 0.0, True, _lib.GRAD_REFERENCE, st2, c2, (0, N, 0))
    hv_pre, hq_pre = (he, hf) if alg == "mc" else (hf, he)
    gB = D + 1                                    # padded layers [0, D + 1) are unaffected by the cut at n_sub
    nv_head, nq_head = int(v_pre[gB]), int(q_pre[gB])
    assert int(hv_pre[gB]) == nv_head and int(hq_pre[gB]) == nq_head
    assert torch.equal(verts[:nv_head], hv[:nv_head]), "head vertices differ"
    assert torch.equal(faces[:nq_head], hfaces[:nq_head]), "head faces differ"
    _log(dict(test="single_call_1024", alg=alg, verts=V, faces=Fn, face_elements=int(faces.numel()), k=k,
              tail_verts=nv_tail, tail_faces=nq_tail, head_verts=nv_head, head_faces=nq_head))
