#!/usr/bin/env python3
"""Per-source-line memory picture of one kernel in an ncu report (CUDA-C view):
   python tools/line_mem.py rep.ncu-rep <kernel-regex> [file-substring]
Columns: warp instructions, L1 tag requests (global), L2 theoretical sectors (global), shared wavefronts (excess)."""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
want = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", "regex:" + rx, "-c", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = cur = None
out = []
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) > 8 and r[0].isdigit():
        def f(k, r=r):
            try:
                return float(r[hdr.index(k) - len(hdr)] or 0)
            except ValueError:
                return 0.0
        out.append((cur, int(r[0]), f("Instructions Executed"), f("L1 Tag Requests Global"), f("L2 Theoretical Sectors Global"),
                    f("L1 Wavefronts Shared"), f("L1 Wavefronts Shared Excessive"), f("# Samples"), r[1].strip()[:90]))
tot = [sum(o[i] for o in out) or 1 for i in range(2, 8)]
print("totals: inst %.4g  L1tag %.4g  L2sect %.4g  shWF %.4g  shExc %.4g samples %.4g" % tuple(tot))
for o in out:
    if want and want not in (o[0] or ""):
        continue
    if o[3] or o[5] or o[2] > 0.004 * tot[0]:
        print("%-26s %4d inst %5.2f%% tag %5.2f%% l2s %5.2f%% shwf %5.2f%% (exc %5.2f%%) smp %5.2f%% | %s" % (
            o[0], o[1], 100 * o[2] / tot[0], 100 * o[3] / tot[1], 100 * o[4] / tot[2], 100 * o[5] / tot[3], 100 * o[6] / tot[3], 100 * o[7] / tot[5], o[8]))
