"""Activation-checkpoint and re-entrancy semantics (the reference "supports checkpointing"
by recomputing the forward inside backward and keeps mutable scratch in the extractor object).  Here nothing
lives in the module: the rank structure of a forward travels with its autograd node, so
  * torch.utils.checkpoint reproduces the plain gradients bit for bit (the backward is deterministic), and
  * forwards of OTHER inputs between a forward and its backward do not disturb it."""
import pytest
import torch
from torch.utils.checkpoint import checkpoint

from diso_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _loss(v):
    w = torch.cos(torch.arange(v.numel(), dtype=torch.float64, device=v.device).reshape(-1, 3) * 0.618).to(v.dtype)
    return (v * w).sum()


@pytest.mark.parametrize("alg", ["mc", "dmc"])
def test_checkpoint_and_interleaved_forwards(alg):
    import diso_b200
    m = diso_b200.DiffMC() if alg == "mc" else diso_b200.DiffDMC()
    kw = {} if alg == "mc" else dict(return_quads=True)
    shape = (30, 26, 41)
    sdf = syn.random_sdf(shape, "flexi", 31).to(DEV)
    deform = (syn.random_deform(shape, 32) * 0.5).to(DEV)
    other = syn.random_sdf(shape, "dense", 33).to(DEV)

    def run(fn):
        s = sdf.clone().requires_grad_(True)
        d = deform.clone().requires_grad_(True)
        fn(s, d)
        return s.grad.clone(), d.grad.clone()

    def plain(s, d):
        _loss(m(s, d, **kw)[0]).backward()

    def checkpointed(s, d):
        v = checkpoint(lambda a, b: m(a, b, **kw)[0], s, d, use_reentrant=False)
        _loss(v).backward()

    def interleaved(s, d):
        v = m(s, d, **kw)[0]
        loss = _loss(v)
        for _ in range(2):                       # the same module extracts other surfaces before the backward runs
            m(other, None, **kw)
            m(other.clone().requires_grad_(True), deform, **kw)
        loss.backward()

    gs, gd = run(plain)
    assert float(gs.abs().sum()) > 0 and float(gd.abs().sum()) > 0
    for fn in (checkpointed, interleaved):
        hs, hd = run(fn)
        assert torch.equal(gs, hs) and torch.equal(gd, hd), fn.__name__
