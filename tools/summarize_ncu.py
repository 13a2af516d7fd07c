#!/usr/bin/env python3
"""Summarise an ncu report (.ncu-rep, --set full) into a markdown table for profiles/.

usage: python tools/summarize_ncu.py gpurun_out/prof_X.ncu-rep profiles/X_summary.md --size 256 [--note "..."]
Runs here on the CPU box (ncu -i ... --page raw --csv)."""
import argparse
import csv
import io
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("out")
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--note", default="")
a = ap.parse_args()

raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, data = rows[0], rows[2:]
n = a.size + 2
nch = n * n * ((n + 31) // 32)


def g(r, name):
    try:
        return float(r[h.index(name)].replace(",", ""))
    except Exception:
        return float("nan")


M = [("us", "gpu__time_duration.sum", 1.0), ("inst/chunk", "smsp__inst_executed.sum", 1.0 / nch),
     ("thr/inst", "smsp__thread_inst_executed_per_inst_executed.ratio", 1.0),
     ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
     ("L1 %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
     ("LSU wf %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1.0),
     ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
     ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
     ("DRAM rd MB", "dram__bytes_read.sum", 1.0), ("DRAM wr MB", "dram__bytes_write.sum", 1.0),
     ("regs", "launch__registers_per_thread", 1.0), ("warps %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
     ("stall long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1.0),
     ("stall barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 1.0)]
units = rows[1]
out = ["# ncu summary: %s" % a.rep.split("/")[-1], "",
       "Workload: `tools/profile_step.py --size %d` (rand-flexi SDF + deform, fp32, DiffMC + DiffDMC forward+backward), "
       "`ncu --set full --clock-control none`; %d chunks of 32 points. Durations under ncu are serialised / cold-cache: "
       "compare shares, not absolutes." % (a.size, nch), ""]
if a.note:
    out += [a.note, ""]
out.append("| kernel | " + " | ".join(m[0] for m in M) + " |")
out.append("|---|" + "---|" * len(M))
for r in data:
    name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("diso::", "")
    vals = []
    for label, key, sc in M:
        v = g(r, key) * sc
        if "bytes" in key:
            u = units[h.index(key)]
            v = v * {"Mbyte": 1.0, "Gbyte": 1000.0, "Kbyte": 0.001, "byte": 1e-6}.get(u, 1.0)
        if key == "gpu__time_duration.sum":   # ncu picks the unit per column: normalise to microseconds
            u = units[h.index(key)]
            v = v * {"us": 1.0, "usecond": 1.0, "ms": 1000.0, "msecond": 1000.0, "ns": 0.001, "nsecond": 0.001, "s": 1e6, "second": 1e6}.get(u, 1.0)
        vals.append("%.1f" % v if abs(v) < 1e5 else "%.3g" % v)
    out.append("| `%s` | " % name[:44] + " | ".join(vals) + " |")
open(a.out, "w").write("\n".join(out) + "\n")
print("\n".join(out))
