#!/usr/bin/env python3
"""SASS listings of the library's kernels for profiles/sass/ (runs on the CPU box: cuobjdump only).

    python tools/dump_sass.py [--lib diso_b200/libdiso_b200.so] [--out profiles/sass]

Writes one <kernel>.sass per fp32 kernel (instruction text only, encodings stripped) and INDEX.md
with the instruction mix (global / shared loads and stores, POPC, barriers, FP64...) of EVERY
function in the library, so a reader can check e.g. that the hot loops use LDG.E.128 / LDS.128,
that no local-memory traffic (LDL/STL) exists, and that nothing went to tensor pipes."""
import argparse
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=os.path.join(ROOT, "diso_b200", "libdiso_b200.so"))
ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "sass"))
a = ap.parse_args()
os.makedirs(a.out, exist_ok=True)
raw = subprocess.run(["cuobjdump", "-sass", a.lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", raw)), capture_output=True, text=True).stdout.split("\n")
blocks = re.split(r"\n\s*Function : ", raw)[1:]
rows = []
for blk, dem in zip(blocks, names):
    body = blk.split("\n", 1)[1]
    short = re.sub(r"\(.*", "", dem).replace("void ", "").replace("diso::", "")
    ins = []
    for line in body.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4})\*/\s+(.*?);?\s*/\*", line)
        if m:
            ins.append((m.group(1), m.group(2).strip()))
    mix = collections.Counter()
    for _, t in ins:
        op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] if t else ""
        base = op.split(".")[0]
        mix[base] += 1
        if base in ("LDG", "STG", "LDS", "STS") and (".128" in op or ".64" in op):
            mix[base + (".128" if ".128" in op else ".64")] += 1
    rows.append((short, len(ins), mix))
    # every fp32 kernel, and the fp64 instantiations of the primary flavours (dense tiles / with deform / unfused ids):
    # BASELINE.json asks for a committed listing of each precision
    is_f32 = "double" not in dem and "debug" not in dem
    is_f64_primary = "double" in dem and "debug" not in dem and not re.search(r"<double, true>|double, \d, true|double, false, \d", short) \
        and "queue" not in short
    if is_f32 or is_f64_primary:
        fn = re.sub(r"[^A-Za-z0-9_]+", "_", short).strip("_") + ".sass"
        with open(os.path.join(a.out, fn), "w") as f:
            f.write("// %s\n// cuobjdump -sass of %s (sm_100a), encodings stripped\n" % (dem, os.path.basename(a.lib)))
            for addr, t in ins:
                f.write("/*%s*/ %s\n" % (addr, t))
cols = ["LDG", "LDG.128", "STG", "STG.128", "LDS", "LDS.128", "STS", "STS.128", "LDL", "STL", "POPC", "SHFL", "BAR", "ATOM", "ATOMS", "RED", "MUFU", "FADD2", "DFMA", "HMMA", "UTCHMMA"]
with open(os.path.join(a.out, "INDEX.md"), "w") as f:
    f.write("# SASS instruction mix per kernel (`tools/dump_sass.py`, static counts)\n\n")
    spill = sorted((short, mix.get("LDL", 0), mix.get("STL", 0)) for short, n, mix in rows if mix.get("LDL", 0) or mix.get("STL", 0))
    f.write("Listings of every fp32 kernel and of the fp64 instantiations of the primary flavours are in this directory (one file\n"
            "per kernel). No tensor pipes (no dense contraction on this path); FADD2 = Blackwell's packed f32x2 add (the accumulators\n"
            "of the saved-record backward).  Local memory: %s.\n\n" % (
                "none" if not spill else "a few spilled words under the 32-register caps -- " + ", ".join("`%s` LDL %d / STL %d" % t for t in spill)))
    f.write("| kernel | instr | " + " | ".join(cols) + " |\n|---|---|" + "---|" * len(cols) + "\n")
    for short, n, mix in sorted(rows):
        f.write("| `%s` | %d | " % (short, n) + " | ".join(str(mix.get(c, 0)) for c in cols) + " |\n")
print("wrote", len(rows), "kernels to", a.out)
