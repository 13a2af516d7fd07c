"""The C-ABI library loads on a GPU-less box and exports exactly what include/diso_b200.h
declares; argument errors are reported through return codes (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from diso_b200 import _build, _lib
    _build.build()
    return _lib.load()


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "diso_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(diso_b200_\w+)\s*\(", hdr)))


def test_header_symbols_are_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "diso_b200", "libdiso_b200.so")],
                         capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (diso_b200_\w+)", out))
    assert exported == set(names), (exported ^ set(names))


def test_ctypes_binding_covers_header():
    from diso_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_library_has_sm100a_code_and_no_torch_dependency():
    so = os.path.join(ROOT, "diso_b200", "libdiso_b200.so")
    elf = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    ldd = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "python" not in ldd


def test_abi_version_and_state_bytes(lib):
    assert lib.diso_b200_abi_version() == 3
    for alg in (0, 1):
        small = lib.diso_b200_state_bytes(alg, 8, 8, 8)
        big = lib.diso_b200_state_bytes(alg, 512, 512, 512)
        assert 0 < small < big
        # compact rank structure incl. the 2 B/cell word array and the active-chunk lists: MC ~3.2 B/voxel,
        # DMC ~3.7 B/voxel at 512^3; the reference keeps >= 4 B/voxel of int32 scratch plus padded copies of
        # both inputs (16 B/voxel)
        assert big < 4.0 * 512 ** 3
    assert lib.diso_b200_state_bytes(1, 64, 64, 64) > lib.diso_b200_state_bytes(0, 64, 64, 64)


def test_argument_errors_return_codes(lib):
    assert lib.diso_b200_state_bytes(0, 0, 4, 4) == 0
    assert lib.diso_b200_state_bytes(7, 4, 4, 4) == 0
    rc = lib.diso_b200_count(0, None, 0, 4, 4, 4, 0.0, None, 0, None)
    assert rc == -1 and b"null" in lib.diso_b200_last_error()
    rc = lib.diso_b200_count(0, ctypes.c_void_p(256), 5, 4, 4, 4, 0.0, ctypes.c_void_p(256), 0, None)
    assert rc == -1 and b"dtype" in lib.diso_b200_last_error()
    rc = lib.diso_b200_count(0, ctypes.c_void_p(256), 0, 4, 4, 4, 0.0, ctypes.c_void_p(256), 16, None)
    assert rc == -2  # state too small
    rc = lib.diso_b200_count(0, ctypes.c_void_p(256), 0, 2000, 2000, 2000, 0.0, ctypes.c_void_p(256), 1 << 40, None)
    assert rc == -4  # too large for one call
    rc = lib.diso_b200_mc_backward(ctypes.c_void_p(256), ctypes.c_void_p(256), 0, 4, 4, 4, 0.0, ctypes.c_void_p(256), None,
                                   ctypes.c_void_p(256), 1, None, None, 0, ctypes.c_void_p(256), None, None)
    assert rc == -1  # no saved edge records: deform without adj_deform is an error
    rc = lib.diso_b200_mc_backward(ctypes.c_void_p(256), None, 0, 4, 4, 4, 0.0, ctypes.c_void_p(256), None,
                                   ctypes.c_void_p(256), 1, None, ctypes.c_void_p(256), 0, ctypes.c_void_p(256), None, None)
    assert rc == -1 and b"edge_rec_stride" in lib.diso_b200_last_error()
