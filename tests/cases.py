"""Shared parity-test cases: (name, sdf builder, deform?, iso) at sizes the CPU oracle finishes
in well under a second each.  Covers the BASELINE configs' shapes in miniature plus the edge
cases of the domain: ragged / non-cubic / tiny grids, rows that are not a multiple of 32 or 4,
iso != 0, values exactly equal to iso, surfaces touching the boundary, empty surfaces."""
import torch

from diso_b200 import synthetic as syn


def _plane(shape, axis, offset):
    idx = torch.arange(shape[axis], dtype=torch.float32) - offset
    view = [1, 1, 1]
    view[axis] = -1
    return idx.view(view).expand(*shape).contiguous()


def _ints(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(-2, 3, shape, generator=g).float()  # many exact ties with iso=0


CASES = {
    # name: (sdf factory, use deform, iso)
    "sphere32": (lambda: syn.sphere_sdf(32, margin=1 / 32), False, 0.0),
    "sphere64": (lambda: syn.sphere_sdf(64), False, 0.0),                     # BASELINE C1
    "roundcube32_def": (lambda: syn.round_cube_sdf(32), True, 0.0),             # BASELINE C2 in miniature
    "rand_flexi_24": (lambda: syn.random_sdf(24, "flexi", 0), True, 0.0),       # C3/C4 in miniature
    "rand_dense_19": (lambda: syn.random_sdf(19, "dense", 0), True, 0.0),       # worst case, odd size
    "rand_sparse_36": (lambda: syn.random_sdf(36, "sparse", 0), False, 0.0),
    "rand_dense_40x33x70": (lambda: syn.random_sdf((40, 33, 70), "dense", 12), True, 0.0),  # larger; no golden, oracle only
    "ragged_5x9x70": (lambda: syn.random_sdf((5, 9, 70), "dense", 3), True, 0.0),   # 3 chunks per row, Z%4!=0
    "ragged_31x2x30": (lambda: syn.random_sdf((31, 2, 30), "flexi", 4), True, 0.0),  # PZ == 32 exactly
    "ragged_3x4x62": (lambda: syn.random_sdf((3, 4, 62), "dense", 5), False, 0.0),   # PZ == 64 exactly
    "longrow_3x4x644": (lambda: syn.random_sdf((3, 4, 644), "dense", 13), True, 0.0),   # > 512 values per row: 2 load batches
    "longrow_2x3x1100": (lambda: syn.random_sdf((2, 3, 1100), "flexi", 14), False, 0.0),  # 3 batches, NC > 32
    "longrow_1x2x1101": (lambda: syn.random_sdf((1, 2, 1101), "dense", 15), True, 0.0),  # scalar sign pass, NC > 32
    "tall_1x530000x2": (lambda: syn.random_sdf((1, 530000, 2), "flexi", 16), True, 0.0),   # > 65535 backward tiles in y: flat-grid fallback
    "deep_40000x2x3": (lambda: syn.random_sdf((40000, 2, 3), "dense", 17), True, 0.0),     # coordinates beyond 16 bits in x
    "tiny_1x1x1": (lambda: torch.full((1, 1, 1), -0.3), True, 0.0),
    "tiny_2x2x2": (lambda: syn.random_sdf(2, "dense", 6), True, 0.0),
    "thin_1x7x33": (lambda: syn.random_sdf((1, 7, 33), "dense", 7), True, 0.0),
    "iso_0p37": (lambda: syn.random_sdf(14, "dense", 8) + 0.5, True, 0.37),
    "iso_neg": (lambda: syn.sphere_sdf(24, margin=1 / 24), False, -0.05),
    "ties_int": (lambda: _ints((6, 7, 37), 9), True, 0.0),                       # values == iso
    "plane_x": (lambda: _plane((9, 8, 40), 0, 3.5), False, 0.0),
    "plane_z_tie": (lambda: _plane((6, 7, 34), 2, 16.0), False, 0.0),              # a whole layer == iso
    "all_inside_but_one": (lambda: torch.ones(6, 6, 6).index_put((torch.tensor(2), torch.tensor(3), torch.tensor(4)), torch.tensor(-1.0)), True, 0.0),
    "boundary_negative": (lambda: -torch.ones(4, 5, 6) + 1.5 * (syn.random_sdf((4, 5, 6), "dense", 10) > 0).float(), True, 0.0),
}

EMPTY_CASES = {
    "all_above": (lambda: torch.ones(5, 6, 7), 0.0),          # min >= iso
    "all_below": (lambda: -torch.ones(5, 6, 7), 0.0),         # max <= iso
    "max_eq_iso": (lambda: torch.zeros(4, 4, 4).index_put((torch.tensor(1), torch.tensor(1), torch.tensor(1)), torch.tensor(-1.0)), 0.0),
    "all_eq_iso": (lambda: torch.zeros(3, 3, 3), 0.0),
}


def make(name, dtype=torch.float32):
    factory, use_def, iso = CASES[name]
    sdf = factory().to(dtype).contiguous()
    deform = syn.random_deform(tuple(sdf.shape), seed=1, dtype=dtype) if use_def else None
    return sdf, deform, iso
