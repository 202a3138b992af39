#!/usr/bin/env python
"""Phase-level breakdown of an `ncu --set full --import-source on` capture of the step kernels: instruction and
stall-sample shares per algorithm phase (by source line ranges of b2g_dynamics.cuh) plus the headline raw metrics.
Usage: python tools/ncu_phases.py gpurun_out/prof.ncu-rep [out.csv]"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
dyn = open("isaacgymenv_b200/csrc/b2g_dynamics.cuh").read().splitlines()
marks = [(i + 1, m.group(1).strip()) for i, l in enumerate(dyn) for m in [re.match(r"\s*// -{8,} (.*?) -{4,}", l)] if m]
marks = [(1, "prologue")] + marks


def phase_of(line):
    name = marks[0][1]
    for ln, nm in marks:
        if line >= ln:
            name = nm
    return name


hdr, cur_file, cur_line, sass = None, None, None, []
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if not hdr or not r:
        continue
    if r[0].isdigit():
        cur_line = int(r[0])
        continue
    if r[0] == "" and len(r) > 6 and r[2].startswith("0x"):
        d = dict(zip(hdr, r))
        sass.append((int(r[2], 16), cur_file, cur_line, int(d["Instructions Executed"] or 0), int(d["# Samples"] or 0)))
sass.sort()
agg, last = {}, "task/kernel (b2g_threads.cuh, b200gym.cu)"
for addr, f, l, i, sm in sass:
    if f == "b2g_dynamics.cuh":
        last = phase_of(l)
    elif f in ("b2g_threads.cuh", "b200gym.cu"):
        last = "task/kernel (b2g_threads.cuh, b200gym.cu)"
    a = agg.setdefault(last, [0, 0, 0])
    a[0] += i
    a[1] += sm
    a[2] += 1
ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
out = [("phase", "sass_instructions", "inst_executed_pct", "stall_samples_pct")]
for k, v in agg.items():
    out.append((k, v[2], round(100 * v[0] / ti, 2), round(100 * v[1] / max(ts, 1), 2)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u, data = rr[0], rr[1], rr[2:]
for key in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_inst0.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "smsp__average_warp_latency_per_inst_issued.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"):
    if key in h:
        i = h.index(key)
        out.append((key, u[i], " ".join(row[i] for row in data), ""))
w = csv.writer(open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout)
w.writerows(out)
