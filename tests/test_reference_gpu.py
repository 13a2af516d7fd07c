"""Side-by-side parity with the UNMODIFIED reference CUDA build (baseline/_ref) on the same GPU,
same inputs -- the check BASELINE.json's north_star asks for.  Skips cleanly when the reference
build is not present in the snapshot."""
import numpy as np
import pytest
import torch

from diso_b200 import synthetic as syn
from tests import cases
from tests.refload import load_reference
from tests.test_parity_gpu import DEV, TOL, close, weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    m = load_reference()
    if m is None:
        pytest.skip("reference build baseline/_ref not available")
    return m


def run(mod, sdf, deform, iso, normalize, **kw):
    s = sdf.to(DEV).requires_grad_(True)
    d = deform.to(DEV).requires_grad_(True) if deform is not None else None
    verts, faces = mod(s, d, isovalue=iso, normalize=normalize, **kw)
    out = dict(verts=verts.detach(), faces=faces)
    if verts.shape[0]:
        (verts * weights(verts.shape[0], verts.dtype, DEV)).sum().backward()
        out["gsdf"] = s.grad
        out["gdef"] = d.grad if d is not None else None
    return out


BIG = {
    "rand_flexi_128": (lambda dt: syn.random_sdf(128, "flexi", 0, dt), True),
    "rand_dense_96": (lambda dt: syn.random_sdf(96, "dense", 0, dt), True),
    "rand_sparse_128": (lambda dt: syn.random_sdf(128, "sparse", 0, dt), False),
    "roundcube128": (lambda dt: syn.round_cube_sdf(128, dtype=dt), True),   # BASELINE C2 at full size
    "sphere64": (lambda dt: syn.sphere_sdf(64, dtype=dt), False),           # BASELINE C1
}


def _inputs(name, dtype):
    if name in BIG:
        f, use_def = BIG[name]
        sdf = f(dtype)
        return sdf, (syn.random_deform(tuple(sdf.shape), 1, dtype) if use_def else None), 0.0
    return cases.make(name, dtype)


NAMES = list(BIG) + ["rand_dense_19", "ragged_5x9x70", "ties_int", "iso_0p37", "thin_1x7x33", "boundary_negative", "tiny_2x2x2"]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("alg", ["mc", "dmc"])
@pytest.mark.parametrize("name", NAMES)
def test_side_by_side(ref, name, alg, dtype):
    import diso_b200
    sdf, deform, iso = _inputs(name, dtype)
    if alg == "mc":
        ours, theirs, kw = diso_b200.DiffMC(dtype), ref.DiffMC(dtype), {}
    else:
        ours, theirs, kw = diso_b200.DiffDMC(dtype), ref.DiffDMC(dtype), dict(return_quads=True)
    for normalize in (True, False):
        a = run(ours, sdf, deform, iso, normalize, **kw)
        b = run(theirs, sdf, deform, iso, normalize, **kw)
        assert a["faces"].dtype == b["faces"].dtype
        assert torch.equal(a["faces"], b["faces"]), "face connectivity differs from the reference"
        va, vb = a["verts"].cpu().numpy(), b["verts"].cpu().numpy()
        # north-star tolerance, absolute on the normalised / lattice frame
        close(va, vb, dtype, "verts")
        if "gsdf" in b:
            close(a["gsdf"].cpu().numpy(), b["gsdf"].cpu().numpy(), dtype, "adj_sdf")
            if deform is not None:
                close(a["gdef"].cpu().numpy(), b["gdef"].cpu().numpy(), dtype, "adj_deform")


def test_vertices_bit_identical_to_reference(ref):
    """Stricter than required: with the arithmetic order pinned (edge_math.cuh) the vertices
    should be bit-identical to the reference's, not merely within 1e-5."""
    import diso_b200
    bad = {}
    for name in ("rand_flexi_128", "roundcube128", "ties_int"):
        for alg in ("mc", "dmc"):
            sdf, deform, iso = _inputs(name, torch.float32)
            kw = {} if alg == "mc" else dict(return_quads=True)
            ours = (diso_b200.DiffMC() if alg == "mc" else diso_b200.DiffDMC())
            theirs = (ref.DiffMC() if alg == "mc" else ref.DiffDMC())
            d = deform.to(DEV) if deform is not None else None
            va, _ = ours(sdf.to(DEV), d, iso, **kw)
            vb, _ = theirs(sdf.to(DEV), d, iso, **kw)
            n = int((va != vb).any(1).sum())
            if n:
                bad["%s/%s" % (name, alg)] = (n, float((va - vb).abs().max()))
    assert not bad, "vertices differ in the last bits: %s" % bad


def test_dmc_triangle_split_matches_reference(ref):
    import diso_b200
    for name in ("roundcube128", "rand_flexi_128", "sphere64"):
        sdf, deform, iso = _inputs(name, torch.float32)
        d = deform.to(DEV) if deform is not None else None
        va, fa = diso_b200.DiffDMC()(sdf.to(DEV), d, iso)
        vb, fb = ref.DiffDMC()(sdf.to(DEV), d, iso)
        assert fa.shape == fb.shape and fa.dtype == fb.dtype and torch.equal(va, vb)
        # the kernel reproduces torch's reduction order (quad_split.cuh): identical on every quad, ties included
        assert torch.equal(fa, fb), "%s: triangle list differs from the reference in %d rows" % (name, int((fa != fb).any(1).sum()))
