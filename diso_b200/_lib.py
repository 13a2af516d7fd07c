"""ctypes binding of the C ABI declared in include/diso_b200.h.

This is the stub a maintainer of the reference would write in place of ``from . import _C``
(/root/reference/diso/__init__.py:6).  There is NO fallback: if the shared library is missing
or cannot be loaded the import of the operators fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# $DISO_B200_LIB selects another build of the same library (kernel experiments: A/B variants side by side)
LIB_PATH = os.environ.get("DISO_B200_LIB") or os.path.join(_HERE, "libdiso_b200.so")

ALG_MC, ALG_DMC = 0, 1
F32, F64 = 0, 1
GRAD_REFERENCE, GRAD_EXACT = 0, 1
COUNT_SLOTS = 8
CNT_VERTS, CNT_FACES, CNT_ANY_GT, CNT_EDGES, CNT_USED, CNT_EDGE_CHUNKS, CNT_CELL_CHUNKS = 0, 1, 2, 3, 4, 5, 6


class Frame(ctypes.Structure):
    """diso_b200_frame (include/diso_b200.h): a slab's place inside a larger grid."""
    _fields_ = [("x_origin", ctypes.c_int32), ("X_global", ctypes.c_int32), ("id_offset", ctypes.c_int64)]


def frame_ptr(frame):
    """None or a Frame -> the void* the C ABI takes (keep the Frame alive while the call runs)."""
    return None if frame is None else ctypes.cast(ctypes.pointer(frame), ctypes.c_void_p)


def counts_array(counts):
    """ctypes int64[COUNT_SLOTS] holding the host copy of the count block (passed to emit / backward)."""
    return (ctypes.c_int64 * COUNT_SLOTS)(*[int(c) for c in counts])

# every symbol include/diso_b200.h declares: name -> (restype, argtypes)
_vp, _i, _d, _sz, _i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_size_t, ctypes.c_int64
SIGNATURES = {
    "diso_b200_abi_version": (_i, []),
    "diso_b200_last_error": (ctypes.c_char_p, []),
    "diso_b200_state_bytes": (_sz, [_i, _i, _i, _i]),
    "diso_b200_state_layout": (_i, [_i, _i, _i, _i, ctypes.POINTER(ctypes.c_int64)]),
    "diso_b200_count": (_i, [_i, _vp, _i, _i, _i, _i, _d, _vp, _sz, _vp]),
    "diso_b200_read_counts": (_i, [_vp, _vp, _vp]),
    "diso_b200_mc_emit": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i64, _vp]),
    "diso_b200_dmc_emit": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "diso_b200_mc_backward": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _vp, _vp, _vp, _i, _vp, _vp, _i64, _vp, _vp, _vp]),
    "diso_b200_dmc_backward": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _vp, _vp, _vp, _i, _vp, _i, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "diso_b200_quad_split_scratch_bytes": (_sz, [_i64]),
    "diso_b200_quad_split": (_i, [_vp, _i, _vp, _i64, _vp, _vp, _vp, _vp]),
    "diso_b200_debug_cell_codes": (_i, [_i, _i, _i, _i, _vp, _vp, _vp]),
    "diso_b200_launch_count": (ctypes.c_longlong, []),
    "diso_b200_profile_enable": (_i, [_i]),
    "diso_b200_profile_dump": (_i, [ctypes.c_char_p, _sz]),
}

_lib = None


class DisoB200Error(RuntimeError):
    pass


def load():
    """Load libdiso_b200.so (once).  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DisoB200Error(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `python -m diso_b200._build` (requires nvcc); there is no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.diso_b200_abi_version() != 3:
            raise DisoB200Error("libdiso_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = load().diso_b200_last_error().decode("utf-8", "replace")
        raise DisoB200Error("libdiso_b200 error %d: %s" % (rc, msg))


def launch_count():
    return int(load().diso_b200_launch_count())


class kernel_profile:
    """Context manager: per-kernel device times (ms) of the library calls made inside it, measured
    with CUDA events on the launching stream.  ``.times`` maps kernel name -> list of ms."""

    def __enter__(self):
        load().diso_b200_profile_enable(1)
        self.times = {}
        return self

    def __exit__(self, *exc):
        L = load()
        buf = ctypes.create_string_buffer(1 << 20)
        rc = L.diso_b200_profile_dump(buf, len(buf))
        L.diso_b200_profile_enable(0)
        check(rc)
        for line in buf.value.decode().splitlines():
            name, ms = line.rsplit(" ", 1)
            self.times.setdefault(name, []).append(float(ms))
        return False
