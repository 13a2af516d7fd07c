"""Randomised parity sweep: 40 seeded random configurations (shape incl. degenerate and non-multiple-of-32 rows, density,
isovalue, deform on/off, dtype, normalize, algorithm) through the public modules against the CPU oracle -- faces and
vertices bit for bit, gradients within the north-star bar.  Complements the hand-picked cases of tests/cases.py."""
import numpy as np
import pytest
import torch

from tests.test_parity_gpu import DEV, close, weights

pytestmark = pytest.mark.gpu


def _config(seed):
    r = np.random.RandomState(1000 + seed)
    dims = [int(r.choice([1, 2, 3, 5, 8, 13, 21, 31, 32, 33, 47, 64, 65, 97])) for _ in range(3)]
    while dims[0] * dims[1] * dims[2] > 60000:       # keep the single-threaded oracle fast
        dims[int(np.argmax(dims))] //= 2
    kind = r.choice(["uniform", "smooth", "ints"])
    iso = float(r.choice([0.0, 0.0, 0.25, -0.3]))
    g = torch.Generator().manual_seed(int(seed))
    if kind == "uniform":
        sdf = torch.rand(dims, generator=g) * 2 - 1 + float(r.uniform(-0.6, 0.6))
    elif kind == "smooth":
        x, y, z = torch.meshgrid(*[torch.linspace(-1, 1, d) if d > 1 else torch.zeros(1) for d in dims], indexing="ij")
        sdf = (x * x + 0.7 * y * y + 1.3 * z * z).sqrt() - float(r.uniform(0.3, 1.1)) + 0.02 * torch.rand(dims, generator=g)
    else:
        sdf = torch.randint(-2, 3, dims, generator=g).float() + iso      # many exact ties with iso
    use_def = bool(r.rand() < 0.7)
    deform = (0.45 * torch.tanh(torch.randn(dims + [3], generator=g))) if use_def else None
    dtype = torch.float64 if r.rand() < 0.3 else torch.float32
    return sdf.to(dtype).contiguous(), (deform.to(dtype).contiguous() if use_def else None), iso, bool(r.rand() < 0.5), dtype


@pytest.mark.parametrize("seed", range(40))
def test_random_configuration_vs_oracle(oracle, seed):
    import diso_b200
    sdf, deform, iso, normalize, dtype = _config(seed)
    sn, dn = sdf.numpy(), (deform.numpy() if deform is not None else None)
    for alg in ("mc", "dmc"):
        mod = diso_b200.DiffMC(dtype) if alg == "mc" else diso_b200.DiffDMC(dtype, grad_mode="exact" if seed % 3 == 0 else "reference")
        s = sdf.to(DEV).requires_grad_(True)
        d = deform.to(DEV).requires_grad_(True) if deform is not None else None
        kw = {} if alg == "mc" else dict(return_quads=True)
        verts, faces = mod(s, d, isovalue=iso, normalize=normalize, **kw)
        ev, ef = oracle.forward(alg, sn, dn, iso, normalize)
        assert tuple(verts.shape) == ev.shape and tuple(faces.shape) == ef.shape, (seed, alg, verts.shape, ev.shape)
        if ev.shape[0] == 0:
            continue
        assert np.array_equal(faces.cpu().numpy(), ef), (seed, alg, "faces")
        assert np.array_equal(verts.detach().cpu().numpy(), ev, equal_nan=True), (seed, alg, "verts")
        w = weights(ev.shape[0], dtype)
        (verts * w.to(DEV)).sum().backward()
        gs, gd = oracle.backward(alg, sn, dn, iso, normalize, w.numpy(), "exact" if (alg == "dmc" and seed % 3 == 0) else "reference")
        close(s.grad.cpu().numpy(), gs, dtype, "adj_sdf seed %d %s" % (seed, alg))
        if dn is not None:
            close(d.grad.cpu().numpy(), gd, dtype, "adj_deform seed %d %s" % (seed, alg))
        if alg == "dmc" and np.isfinite(ev).all():      # default path: triangle list of the numpy restatement (torch's reduction order)
            vt, ft = mod(sdf.to(DEV), deform.to(DEV) if deform is not None else None, isovalue=iso, normalize=normalize)
            et, _, _ = oracle.split_quads(ev, ef)
            assert np.array_equal(ft.cpu().numpy(), et), (seed, "triangle split")
