"""Builds libdiso_b200.so in-tree with nvcc for sm_100a (no torch / pybind dependency).

Used by ``__graft_entry__.build()``; can also be run directly: ``python -m diso_b200._build``.
nvcc cross-compiles without a GPU.  The shared object is git-ignored but travels to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdiso_b200.so")
SOURCES = ["api.cu"]
DEPS = ["api.cu", "common.cuh", "classify.cuh", "edge_math.cuh", "mc_backward_compact.cuh", "mc_backward_v2.cuh", "compact.cuh", "dmc_compact.cuh", "quad_split.cuh",
        "tables.cuh", "case_tables.inc", os.path.join("..", "..", "include", "diso_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # The reference is built with nvcc defaults (-fmad=true).  We disable automatic contraction
    # and spell the one FMA of the reference's vertex expression explicitly (edge_math.cuh), so
    # the arithmetic is pinned by the source rather than by the optimiser.
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set $NVCC)")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for d in DEPS:
        p = os.path.join(CSRC, d)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build(force=False, verbose=False, out=None, extra=()):
    """Compile the library.  `out` / `extra` build an experimental variant beside the product
    (e.g. out=".../libdiso_b200_x.so", extra=["-DDISO_TUNE"]); select it with $DISO_B200_LIB."""
    if out is None and not force and not needs_build():
        return LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out or LIB] + SOURCES
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        sys.stderr.write(r.stderr)
    return out or LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
