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
    is_f32 = "double" not in dem and "debug" not in dem
    if is_f32:
        fn = re.sub(r"[^A-Za-z0-9_]+", "_", short).strip("_") + ".sass"
        with open(os.path.join(a.out, fn), "w") as f:
            f.write("// %s\n// cuobjdump -sass of %s (sm_100a), encodings stripped\n" % (dem, os.path.basename(a.lib)))
            for addr, t in ins:
                f.write("/*%s*/ %s\n" % (addr, t))
cols = ["LDG", "LDG.128", "STG", "STG.128", "LDS", "LDS.128", "STS", "STS.128", "LDL", "STL", "POPC", "SHFL", "BAR", "ATOM", "ATOMS", "RED", "MUFU", "DFMA", "HMMA", "UTCHMMA"]
with open(os.path.join(a.out, "INDEX.md"), "w") as f:
    f.write("# SASS instruction mix per kernel (`tools/dump_sass.py`, static counts)\n\n")
    f.write("Listings of the fp32 kernels are in this directory (one file per kernel). No kernel uses local memory\n(LDL/STL = 0) or tensor pipes (no dense contraction on this path).\n\n")
    f.write("| kernel | instr | " + " | ".join(cols) + " |\n|---|---|" + "---|" * len(cols) + "\n")
    for short, n, mix in sorted(rows):
        f.write("| `%s` | %d | " % (short, n) + " | ".join(str(mix.get(c, 0)) for c in cols) + " |\n")
print("wrote", len(rows), "kernels to", a.out)
