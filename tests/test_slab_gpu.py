"""Slab sharding through the CUDA library (diso_b200/parallel.py, fused path): W gloo ranks share
cuda:0, each extracts its slab with a frame (include/diso_b200.h: diso_b200_frame); the rank-order
concatenation must equal the single-GPU extraction of the whole grid -- connectivity AND vertex
bits (the kernels form the global integer coordinate before adding the deformation), gradients
within the fp32 tolerance."""
import numpy as np
import pytest
import torch

from diso_b200 import synthetic as syn
from tests.test_slab_sharding import _run_sharded, _weights

pytestmark = pytest.mark.gpu

CASES = [("mc", 2, "flexi", True), ("dmc", 2, "flexi", True), ("mc", 3, "dense", False), ("dmc", 3, "sphere", True), ("mc", 2, "lowhalf", True)]


@pytest.mark.parametrize("field", [False, True], ids=["own", "field"])
@pytest.mark.parametrize("alg,world,kind,use_def", CASES)
def test_sharded_cuda_equals_single_gpu(tmp_path, alg, world, kind, use_def, field):
    """field=False: extract_slab on the rank's own layers; field=True: the SlabField API (extended leaf, in-place halo
    refresh, gradients of the halo layers returned to their owners inside backward)."""
    import diso_b200
    shape = (23, 18, 37)
    if kind == "sphere":
        sdf = syn.sphere_sdf(shape[0])[:, :shape[1], :shape[2]].contiguous()
    elif kind == "lowhalf":       # the last slab owns nothing: everything above x = 9 is outside
        sdf = syn.random_sdf(shape, "flexi", 5)
        sdf[10:] = 2.0
    else:
        sdf = syn.random_sdf(shape, kind, 11)
    deform = syn.random_deform(tuple(sdf.shape), 12) * 0.5 if use_def else None
    verts, faces, gsdf, gdef = _run_sharded(world, alg, sdf, deform, 0.0, True, True, tmp_path, field)

    s = sdf.cuda().requires_grad_(True)
    d = deform.cuda().requires_grad_(True) if use_def else None
    m = diso_b200.DiffMC() if alg == "mc" else diso_b200.DiffDMC()
    kw = {} if alg == "mc" else dict(return_quads=True)
    ev, ef = m(s, d, **kw)
    (ev * _weights(ev.shape[0], torch.float32).cuda()).sum().backward()
    assert torch.equal(faces, ef.cpu()), "connectivity of the stitched mesh differs"
    assert torch.equal(verts, ev.detach().cpu()), "stitched vertices are not bit-identical to the single-GPU run"
    gs = s.grad.cpu().numpy()
    np.testing.assert_allclose(gsdf.numpy(), gs, rtol=0, atol=1e-5 * max(1.0, np.abs(gs).max()))
    if use_def:
        gd = d.grad.cpu().numpy()
        np.testing.assert_allclose(gdef.numpy(), gd, rtol=0, atol=1e-5 * max(1.0, np.abs(gd).max()))


@pytest.mark.parametrize("world", [2])
def test_slab_over_nccl(world):
    """The same check over the NCCL backend (one process per GPU, halo P2P + all_gather over NVLink): needs >= 2 GPUs,
    so it runs in the multi-GPU tiers and skips on a single-GPU box."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(root, "tools", "slab_nccl_check.py"), "--size", "96"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "faces_equal=True" in r.stdout
