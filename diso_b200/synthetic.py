"""Seeded synthetic SDF inputs (BASELINE.md section 2 / SURVEY.md section 8d).

Everything is generated on the CPU with an explicit ``torch.Generator`` so that every
implementation (reference CUDA build, CPU restatement, this library) and every GPU sees identical
bits; callers move the tensors to the device.
"""
import torch


def sphere_sdf(n=64, radius=0.5, margin=1.0 / 64, dtype=torch.float32):
    """The sphere of the reference's smoke script (test/example.py:9-56)."""
    c = torch.zeros(3)
    lo, hi = c - radius - margin, c + radius + margin
    ax = [torch.linspace(0, 1, n) for _ in range(3)]
    g = torch.stack(torch.meshgrid(*ax, indexing="ij"), dim=-1)
    for i in range(3):
        g[..., i] = g[..., i] * (hi[i] - lo[i]) + lo[i]
    return (torch.norm(g - c, dim=-1) - radius).to(dtype).contiguous()


def round_cube_sdf(n=128, half=0.35, r=0.1, dtype=torch.float32):
    """Rounded box (survey-defined C2): |max(|p|-h,0)| + min(max(|p|-h),0) - r on [-.5,.5]^3."""
    ax = torch.linspace(-0.5, 0.5, n)
    p = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), dim=-1)
    q = p.abs() - half
    sdf = torch.norm(q.clamp(min=0), dim=-1) + q.max(dim=-1).values.clamp(max=0) - r
    return sdf.to(dtype).contiguous()


RANDOM_OFFSETS = {"flexi": 0.1, "sparse": 0.0266, "dense": 0.5}


def random_sdf(shape, kind="flexi", seed=0, dtype=torch.float32):
    """U(0,1) - offset: 'flexi' (matches the README's rand-init ratios), 'sparse', 'dense'."""
    if isinstance(shape, int):
        shape = (shape,) * 3
    gen = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.rand(*shape, generator=gen, dtype=torch.float32) - RANDOM_OFFSETS[kind]).to(dtype).contiguous()


def random_deform(shape, seed=1, dtype=torch.float32):
    """0.5 * tanh(U(0,1)) per test/example.py:59-70 (keeps the lattice fold-free)."""
    if isinstance(shape, int):
        shape = (shape,) * 3
    gen = torch.Generator(device="cpu").manual_seed(seed)
    return (0.5 * torch.tanh(torch.rand(*shape, 3, generator=gen, dtype=torch.float32))).to(dtype).contiguous()
