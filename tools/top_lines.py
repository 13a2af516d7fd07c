#!/usr/bin/env python3
"""Top source lines of one kernel in an ncu report: python tools/top_lines.py rep.ncu-rep <kernel-regex> [N]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", "regex:" + rx, "-c", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = cur = None
out = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) > 8 and r[0].isdigit():
        def f(k, r=r):  # index from the END: source text with quotes/commas can add leading columns
            try:
                return float(r[hdr.index(k) - len(hdr)] or 0)
            except ValueError:
                return 0.0
        out.append((f("Instructions Executed"), f("# Samples"), f("L1 Tag Requests Global"), f("L1 Wavefronts Shared"), cur, int(r[0]), r[1].strip()[:95]))
ti, ts, tg, tsh = (sum(o[i] for o in out) or 1 for i in range(4))
print("inst %.3g samples %.3g L1-global-req %.3g smem-wavefronts %.3g" % (ti, ts, tg, tsh))
for o in sorted(out, reverse=True)[:N]:
    print("%5.1f%% inst %5.1f%% smp %5.1f%% gl %5.1f%% sh  %s:%d  %s" % (100 * o[0] / ti, 100 * o[1] / ts, 100 * o[2] / tg, 100 * o[3] / tsh, o[4], o[5], o[6]))
