"""The quad-split kernel (diso_b200/csrc/quad_split.cuh) evaluates the reference's 12 triangle-angle cosines
(diso/__init__.py:118-147: 4 triangles x 3 angles, each from two freshly normalised edge vectors) from SIX unit
vectors and sign flips.  That is bit-exact because IEEE subtraction, division and the dot product's
multiply / add sequence are sign-symmetric.  Checked here in numpy (IEEE fp32 / fp64, same operation order as
the kernel: no fused multiply-add in the dot product, fma only inside the norm) on random and degenerate quads."""
import numpy as np
import pytest


def _fma(a, b, c):
    # exact a*b+c rounded once, via extended precision (float32) / Dekker-free shortcut for the test sizes (float64: longdouble)
    wide = np.float64 if a.dtype == np.float32 else np.longdouble
    return (a.astype(wide) * b.astype(wide) + c.astype(wide)).astype(a.dtype)


def unit(a, b):
    v = a - b
    n2 = v[..., 0] * v[..., 0]
    n2 = _fma(v[..., 1], v[..., 1], n2)
    n2 = _fma(v[..., 2], v[..., 2], n2)
    n = np.sqrt(n2)
    n = np.where(n > v.dtype.type(1e-12), n, v.dtype.type(1e-12))
    return v / n[..., None]


def dot3(a, b):
    s = a[..., 0] * b[..., 0]
    s = s + a[..., 1] * b[..., 1]
    s = s + a[..., 2] * b[..., 2]
    return s


def tri_cos(v0, v1, v2):   # the reference's three cosines of triangle (v0, v1, v2)
    return [dot3(unit(v1, v0), unit(v2, v0)), dot3(unit(v2, v1), unit(v0, v1)), dot3(unit(v0, v2), unit(v1, v2))]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_six_unit_vectors_reproduce_the_twelve_cosines(dtype):
    rng = np.random.default_rng(5)
    q = rng.standard_normal((20000, 4, 3)).astype(dtype)
    q[:500, 1] = q[:500, 0]                      # zero-length edges
    q[500:1000] = np.round(q[500:1000] * 2) / 2   # symmetric / tied configurations
    q[1000:1500, :, 2] = 0                        # planar quads
    v = [q[:, i] for i in range(4)]
    ref = tri_cos(v[0], v[1], v[3]) + tri_cos(v[1], v[2], v[3]) + tri_cos(v[0], v[1], v[2]) + tri_cos(v[0], v[2], v[3])
    S0, S1, S2, S3 = unit(v[1], v[0]), unit(v[2], v[1]), unit(v[3], v[2]), unit(v[0], v[3])
    D0, D1 = unit(v[2], v[0]), unit(v[3], v[1])
    ours = [-dot3(S0, S3), -dot3(D1, S0), -dot3(S3, D1),
            dot3(S1, D1), -dot3(S2, S1), dot3(D1, S2),
            dot3(S0, D0), -dot3(S1, S0), dot3(D0, S1),
            -dot3(D0, S3), -dot3(S2, D0), -dot3(S3, S2)]
    for i, (a, b) in enumerate(zip(ref, ours)):
        assert np.array_equal(a, b), "cosine %d differs in %d of %d quads" % (i, int((a != b).sum()), len(a))
