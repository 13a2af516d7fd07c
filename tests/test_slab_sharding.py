"""Slab sharding of one grid across ranks (diso_b200/parallel.py).

CPU (gloo, world size 2 and 3): the sharding / halo-exchange / id-stitching logic with the CPU
oracle as the per-slab extractor, checked against the oracle run on the whole grid.
GPU: the same with the CUDA extractor (two gloo ranks sharing cuda:0), checked against the
single-GPU operator on the whole grid."""
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from diso_b200 import parallel, synthetic as syn
from tests.slab_helpers import field_worker, worker


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _weights(n, dtype):
    i = torch.arange(n * 3, dtype=torch.float64).reshape(n, 3)
    return torch.cos(i * 0.6180339887 + 0.25).to(dtype)


def _run_sharded(world, alg, sdf, deform, iso, normalize, use_cuda, tmp_path, field=False):
    mp.spawn(worker, args=(world, _port(), alg, sdf, deform, iso, normalize, use_cuda, str(tmp_path), field), nprocs=world, join=True)
    parts = [torch.load(tmp_path / ("rank%d.pt" % r), weights_only=False) for r in range(world)]
    verts = torch.cat([p["verts"] for p in parts])
    faces = torch.cat([p["faces"] for p in parts])
    gsdf = torch.cat([p["gsdf"] for p in parts])
    gdef = torch.cat([p["gdef"] for p in parts]) if deform is not None else None
    for r, p in enumerate(parts):
        assert p["info"]["vert_offset"] == sum(q["verts"].shape[0] for q in parts[:r])
        assert p["info"]["n_verts_total"] == verts.shape[0] and p["info"]["n_faces_total"] == faces.shape[0]
    return verts, faces, gsdf, gdef


def test_plan_slabs():
    assert parallel.plan_slabs(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert parallel.plan_slabs(1024, 8)[-1] == (896, 1024)
    with pytest.raises(ValueError):
        parallel.plan_slabs(3, 2)


@pytest.mark.parametrize("world", [2, 3])
def test_slab_field_halo_refresh_and_gradient_return(tmp_path, world):
    """parallel.SlabField on the CPU (gloo): halos are filled in place from the neighbours, and in backward the gradients
    that land on halo layers are sent to their owners and added there -- every layer of the grid ends up with the
    gradient contributions of ALL ranks that looked at it."""
    X = 11
    full = torch.arange(X * 3 * 4, dtype=torch.float64).reshape(X, 3, 4) / 7.0
    mp.spawn(field_worker, args=(world, _port(), full, str(tmp_path)), nprocs=world, join=True)
    grad = torch.zeros_like(full)
    for r in range(world):
        p = torch.load(tmp_path / ("field%d.pt" % r), weights_only=False)
        grad[p["a"]: p["b"]] = p["grad"]
    # expected: layer x is seen by its owner and by every neighbour whose halo covers it; each contributes 2 v (x + 1)
    slabs = parallel.plan_slabs(X, world)
    seen = torch.zeros(X)
    for r, (a, b) in enumerate(slabs):
        lo = a - (parallel.HALO if r > 0 else 0)
        hi = b + (parallel.HALO if r < world - 1 else 0)
        seen[lo:hi] += 1
    want = 2 * full * (torch.arange(X, dtype=torch.float64) + 1.0).view(-1, 1, 1) * seen.view(-1, 1, 1).double()
    assert torch.allclose(grad, want, rtol=1e-13, atol=0)


CASES = [("mc", 2, True), ("dmc", 2, True), ("mc", 3, False), ("dmc", 3, True)]


@pytest.mark.parametrize("alg,world,use_def", CASES)
def test_sharded_equals_whole_grid_cpu_oracle(oracle, tmp_path, alg, world, use_def):
    sdf = syn.random_sdf((13, 7, 9), "dense", 21)
    deform = syn.random_deform((13, 7, 9), 22) * 0.5 if use_def else None
    verts, faces, gsdf, gdef = _run_sharded(world, alg, sdf, deform, 0.0, True, False, tmp_path)
    ev, ef = oracle.forward(alg, sdf.numpy(), deform.numpy() if use_def else None, 0.0, True)
    assert np.array_equal(faces.numpy(), ef), "connectivity of the stitched mesh differs"
    np.testing.assert_allclose(verts.numpy(), ev, rtol=0, atol=2e-6)
    w = _weights(ev.shape[0], torch.float32).numpy()
    egs, egd = oracle.backward(alg, sdf.numpy(), deform.numpy() if use_def else None, 0.0, True, w, "reference")
    np.testing.assert_allclose(gsdf.numpy(), egs, rtol=0, atol=1e-5 * max(1.0, np.abs(egs).max()))
    if use_def:
        np.testing.assert_allclose(gdef.numpy(), egd, rtol=0, atol=1e-5 * max(1.0, np.abs(egd).max()))


def test_sharded_globally_empty_and_locally_empty_slabs(oracle, tmp_path):
    # slab 0 entirely below iso (locally "max <= iso") but the global grid is not empty
    sdf = torch.cat([-torch.ones(6, 5, 6), syn.random_sdf((6, 5, 6), "dense", 3)])
    verts, faces, _, _ = _run_sharded(2, "mc", sdf, None, 0.0, False, False, tmp_path)
    ev, ef = oracle.forward("mc", sdf.numpy(), None, 0.0, False)
    assert np.array_equal(faces.numpy(), ef)
    np.testing.assert_allclose(verts.numpy(), ev, rtol=0, atol=2e-6)
    verts, faces, _, _ = _run_sharded(2, "dmc", torch.ones(8, 4, 4), None, 0.0, True, False, tmp_path)
    assert verts.shape == (0, 3) and faces.shape == (0, 4)


@pytest.mark.gpu
@pytest.mark.parametrize("alg", ["mc", "dmc"])
def test_sharded_equals_single_gpu(tmp_path, alg):
    import diso_b200
    sdf = syn.random_sdf((48, 40, 70), "flexi", 5)
    deform = syn.random_deform((48, 40, 70), 6)
    verts, faces, gsdf, gdef = _run_sharded(2, alg, sdf, deform, 0.0, True, True, tmp_path)
    s = sdf.cuda().requires_grad_(True)
    d = deform.cuda().requires_grad_(True)
    m = diso_b200.DiffMC() if alg == "mc" else diso_b200.DiffDMC()
    kw = {} if alg == "mc" else dict(return_quads=True)
    ev, ef = m(s, d, **kw)
    assert torch.equal(faces, ef.cpu())
    assert (verts - ev.detach().cpu()).abs().max() <= 2e-6
    (ev * _weights(ev.shape[0], torch.float32).cuda()).sum().backward()
    assert (gsdf - s.grad.cpu()).abs().max() <= 1e-5 * max(1.0, float(s.grad.abs().max()))
    assert (gdef - d.grad.cpu()).abs().max() <= 1e-5 * max(1.0, float(d.grad.abs().max()))
