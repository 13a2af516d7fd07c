"""the `_C` compatibility shim against the reference's own `diso._C`, the
batched entry points, and a trimesh-free port of test/example.py as an integration test (config C1)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from diso_b200 import synthetic as syn
from tests import cases
from tests.refload import load_reference

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _padded(name, dtype=torch.float32):
    sdf, deform, iso = cases.make(name, dtype)
    g = F.pad(sdf, (1, 1, 1, 1, 1, 1), "constant", iso + 1).contiguous().to(DEV)
    d = F.pad(deform, (0, 0, 1, 1, 1, 1, 1, 1), "constant", 0).contiguous().to(DEV) if deform is not None else None
    return g, d, iso


@pytest.mark.parametrize("cls", ["CUMCFloat", "CUDMCFloat", "CUMCDouble", "CUDMCDouble"])
@pytest.mark.parametrize("name", ["rand_dense_19", "ragged_5x9x70", "sphere32"])
def test_C_shim_matches_reference_C(cls, name):
    ref = load_reference()
    if ref is None:
        pytest.skip("reference build baseline/_ref not available")
    from diso_b200 import _C
    dtype = torch.float64 if cls.endswith("Double") else torch.float32
    g, d, iso = _padded(name, dtype)
    ours, theirs = getattr(_C, cls)(), getattr(ref._C, cls)()
    args = (g, iso) if d is None else (g, d, iso)
    va, fa = ours.forward(*args)
    vb, fb = theirs.forward(*args)
    assert fa.dtype == fb.dtype == torch.int32 and torch.equal(fa, fb)
    tol = 1e-6 if dtype == torch.float32 else 1e-13
    assert (va - vb).abs().max() <= tol * max(g.shape)            # 1 ulp of the coordinate
    w = torch.cos(torch.arange(va.numel(), dtype=torch.float64, device=DEV).reshape(-1, 3) * 0.618 + 0.25).to(dtype)
    outs = []
    for obj in (ours, theirs):
        ag = torch.zeros_like(g)
        if d is None:
            obj.backward(g, iso, w, ag)
            outs.append((ag, None))
        else:
            ad = torch.zeros_like(d)
            obj.backward(g, d, iso, w, ag, ad)
            outs.append((ag, ad))
    gt = 1e-5 if dtype == torch.float32 else 1e-12
    assert (outs[0][0] - outs[1][0]).abs().max() <= gt * max(1.0, float(outs[1][0].abs().max()))
    if d is not None:
        assert (outs[0][1] - outs[1][1]).abs().max() <= gt * max(1.0, float(outs[1][1].abs().max()))


def test_C_shim_rejects_bad_inputs():
    from diso_b200 import _C, DisoB200Error
    m = _C.CUMCFloat()
    with pytest.raises(DisoB200Error, match="CUDA"):
        m.forward(torch.zeros(4, 4, 4), 0.0)
    with pytest.raises(DisoB200Error, match="contiguous"):
        m.forward(torch.zeros(4, 4, 8, device=DEV)[:, :, ::2], 0.0)
    with pytest.raises(DisoB200Error, match="type"):
        m.forward(torch.zeros(4, 4, 4, device=DEV, dtype=torch.float64), 0.0)




def test_C_shim_refuses_unpadded_boundary():
    """The reference's _C level adds no pad; a boundary below iso would give a different (open) mesh there, so the shim
    refuses instead of silently closing the surface with its virtual shell."""
    from diso_b200 import _C, DisoB200Error
    g = syn.random_sdf((8, 9, 10), "dense", 3).to(DEV)          # values in (-0.5, 0.5): the boundary crosses iso = 0
    with pytest.raises(DisoB200Error, match="padded input"):
        _C.CUMCFloat().forward(g, 0.0)
    gp = F.pad(g, (1, 1, 1, 1, 1, 1), "constant", 1.0).contiguous()
    v, f = _C.CUMCFloat().forward(gp, 0.0)
    assert v.shape[0] > 0 and f.dtype == torch.int32


@pytest.mark.parametrize("alg", ["mc", "dmc"])
def test_forward_batch_equals_loop(alg):
    import diso_b200
    m = diso_b200.DiffMC() if alg == "mc" else diso_b200.DiffDMC()
    names = ["sphere32", "rand_flexi_24", "tiny_1x1x1", "rand_dense_19"]   # incl. an empty surface
    grids, deforms = [], []
    for n in names:
        s, d, _ = cases.make(n)
        grids.append(s.to(DEV).requires_grad_(True))
        deforms.append(d.to(DEV) if d is not None else None)
    res = m.forward_batch(grids, deforms)
    assert len(res) == len(names)
    loss = 0
    for (v, f), g, d in zip(res, grids, deforms):
        v1, f1 = m(g.detach(), d)
        assert torch.equal(v.detach(), v1) and torch.equal(f, f1) and f.dtype == f1.dtype
        if v.shape[0]:
            loss = loss + (v ** 2).sum()
    loss.backward()
    assert all(g.grad is not None for g, (v, _) in zip(grids, res) if v.shape[0])


def test_example_script_port():
    """test/example.py of the reference without trimesh: sphere 64^3, both extractors, with and
    without deform, L = ||verts||, gradients finite and non-trivial; mesh is a closed 2-manifold."""
    import diso_b200
    sdf = syn.sphere_sdf(64).to(DEV)
    sdf = torch.nn.Parameter(sdf.clone(), requires_grad=True)
    gen = torch.Generator().manual_seed(0)
    deform = torch.nn.Parameter(torch.rand((64, 64, 64, 3), generator=gen).to(DEV), requires_grad=True)
    lo, hi = -0.5 - 1 / 64, 0.5 + 1 / 64
    for mod in (diso_b200.DiffMC(dtype=torch.float32), diso_b200.DiffDMC(dtype=torch.float32)):
        for use_def in (True, False):
            sdf.grad = deform.grad = None
            verts, faces = mod(sdf, 0.5 * torch.tanh(deform) if use_def else None, isovalue=0)
            world = verts * (hi - lo) + lo
            L = torch.norm(world)
            L.backward()
            assert torch.isfinite(sdf.grad).all() and float(sdf.grad.abs().max()) > 0
            if use_def:
                assert torch.isfinite(deform.grad).all() and float(deform.grad.abs().max()) > 0
            else:
                r = world.norm(dim=1)
                assert float((r - 0.5).abs().max()) < 0.02          # vertices lie on the sphere
            f = faces.cpu().numpy()
            e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), 1)
            _, cnt = np.unique(e, axis=0, return_counts=True)
            assert (cnt == 2).all() and f.max() < verts.shape[0]
