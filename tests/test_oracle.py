"""CPU oracle self-checks: known answers from the survey, mesh invariants, and finite-difference
checks of the adjoints (fp64).  These pin the restatement's internal consistency; its agreement
with the real reference is pinned by tests/test_oracle_golden.py."""
import numpy as np
import pytest
import torch

from diso_b200 import synthetic as syn
from tests import cases


def test_sphere64_known_counts(oracle):
    # 17 618 used cells, 17 616 MC verts, 35 228 tris, 17 618 dual verts
    s = syn.sphere_sdf(64).numpy()
    g, _ = oracle.pad_inputs(s, None, 0.0)
    r = oracle.raw_forward("mc", g, None, 0.0)
    assert len(r["used_index"]) == 17618 and r["verts"].shape == (17616, 3) and r["faces"].shape == (35228, 3)
    r = oracle.raw_forward("dmc", g, None, 0.0)
    assert r["verts"].shape == (17618, 3) and r["faces"].shape == (17616, 4)


@pytest.mark.parametrize("kind,exp", [("dense", (1.511, 3.228, 1.437)), ("flexi", (0.541, 0.963, 0.663)), ("sparse", (0.156, 0.229, 0.205))])
def test_random_density_table(oracle, kind, exp):
    s = syn.random_sdf(64, kind, 0).numpy()
    v, f = oracle.forward("mc", s)
    v2, q = oracle.forward("dmc", s)
    G = s.size
    assert len(q) == len(v)  # one quad per crossing edge
    assert abs(len(v) / G - exp[0]) < 0.03 and abs(len(f) / G - exp[1]) < 0.06 and abs(len(v2) / G - exp[2]) < 0.03


def _edge_counts(faces):
    k = faces.shape[1]
    e = np.concatenate([faces[:, [i, (i + 1) % k]] for i in range(k)], 0)
    e = np.sort(e, 1)
    _, cnt = np.unique(e, axis=0, return_counts=True)
    return cnt


@pytest.mark.parametrize("name", ["sphere32", "roundcube32_def", "rand_sparse_36", "iso_neg"])
def test_meshes_are_closed(oracle, name):
    sdf, deform, iso = cases.make(name)
    v, f = oracle.forward("mc", sdf.numpy(), None if deform is None else deform.numpy(), iso)
    assert f.min() >= 0 and f.max() < len(v)
    assert (_edge_counts(f) == 2).all()  # the iso+1 pad closes the surface at the boundary
    v, q = oracle.forward("dmc", sdf.numpy(), None if deform is None else deform.numpy(), iso)
    assert q.min() >= 0 and q.max() < len(v)
    assert (_edge_counts(q) % 2 == 0).all()


def test_empty_early_out(oracle):
    for name, (factory, iso) in cases.EMPTY_CASES.items():
        v, f = oracle.forward("mc", factory().numpy(), None, iso)
        assert v.shape == (0, 3) and f.shape == (0, 3) and f.dtype == np.int32, name
        v, f = oracle.forward("dmc", factory().numpy(), None, iso)
        assert v.shape == (0, 3) and f.shape == (0, 4) and f.dtype == np.int32, name


@pytest.mark.parametrize("alg,mode", [("mc", "reference"), ("dmc", "exact")])
def test_adjoint_matches_finite_differences(oracle, alg, mode):
    rng = np.random.default_rng(0)
    sdf = syn.random_sdf((5, 6, 7), "dense", 11, torch.float64).numpy()
    deform = syn.random_deform((5, 6, 7), 1, torch.float64).numpy() * 0.3
    v0, _ = oracle.forward(alg, sdf, deform, 0.0, True)
    w = rng.standard_normal(v0.shape)
    gs, gd = oracle.backward(alg, sdf, deform, 0.0, True, w, mode)
    eps = 1e-7
    for _ in range(12):
        i = tuple(rng.integers(0, n) for n in sdf.shape)
        sp = sdf.copy(); sp[i] += eps
        sm = sdf.copy(); sm[i] -= eps
        vp, _ = oracle.forward(alg, sp, deform, 0.0, True)
        vm, _ = oracle.forward(alg, sm, deform, 0.0, True)
        assert vp.shape == v0.shape
        fd = ((vp - vm) * w).sum() / (2 * eps)
        assert abs(fd - gs[i]) <= 1e-5 * max(1.0, abs(fd)), (i, fd, gs[i])
        j = i + (int(rng.integers(0, 3)),)
        dp = deform.copy(); dp[j] += eps
        dm = deform.copy(); dm[j] -= eps
        vp, _ = oracle.forward(alg, sdf, dp, 0.0, True)
        vm, _ = oracle.forward(alg, sdf, dm, 0.0, True)
        fd = ((vp - vm) * w).sum() / (2 * eps)
        assert abs(fd - gd[j]) <= 1e-5 * max(1.0, abs(fd)), (j, fd, gd[j])


def test_dmc_reference_grad_mode_differs_only_with_multi_patch_cells(oracle):
    s = syn.sphere_sdf(24, margin=1 / 24, dtype=torch.float64).numpy()  # single-patch cells only
    v, _ = oracle.forward("dmc", s)
    w = np.cos(np.arange(v.size).reshape(v.shape) * 0.618 + 0.25)
    a, _ = oracle.backward("dmc", s, None, 0.0, True, w, "reference")
    b, _ = oracle.backward("dmc", s, None, 0.0, True, w, "exact")
    assert np.array_equal(a, b)
    s = syn.random_sdf(12, "dense", 2, torch.float64).numpy()
    v, _ = oracle.forward("dmc", s)
    w = np.cos(np.arange(v.size).reshape(v.shape) * 0.618 + 0.25)
    a, _ = oracle.backward("dmc", s, None, 0.0, True, w, "reference")
    b, _ = oracle.backward("dmc", s, None, 0.0, True, w, "exact")
    assert not np.allclose(a, b)


def test_deform_gradient_checksum(oracle):
    # every vertex is a convex combination of its edge's endpoints, so in the PADDED frame
    # sum(adj_deform) == sum(adj_verts) (the API-level slice drops the pad layer's share)
    sdf, deform, iso = cases.make("rand_flexi_24", torch.float64)
    g, d = oracle.pad_inputs(sdf.numpy(), deform.numpy(), iso)
    for alg, mode in (("mc", "reference"), ("dmc", "exact")):
        v = oracle.raw_forward(alg, g, d, iso)["verts"]
        w = np.cos(np.arange(v.size).reshape(v.shape) * 0.618 + 0.25)
        _, gd = oracle.raw_backward(alg, g, d, iso, w, mode)
        np.testing.assert_allclose(gd.reshape(-1, 3).sum(0), w.sum(0), rtol=0, atol=1e-9)
