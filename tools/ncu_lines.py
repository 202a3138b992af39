#!/usr/bin/env python
"""Summarise an `ncu --set full --import-source on` capture of one kernel: headline raw metrics plus the source lines with
the most executed instructions / stall samples.   python tools/ncu_lines.py rep.ncu-rep source.cu out.csv [launch_index]"""
import collections
import csv
import subprocess
import sys

rep, srcfile, out = sys.argv[1], sys.argv[2], sys.argv[3]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
lines = open(srcfile).read().splitlines()
base = srcfile.split("/")[-1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr, cur_line, cur_file = None, None, None
per = collections.defaultdict(lambda: [0, 0, 0])
ti = ts = 0
for r in csv.reader(src.splitlines()):
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
        i, s = int(d["Instructions Executed"] or 0), int(d["# Samples"] or 0)
        p = per[(cur_file, cur_line)]
        p[0] += i
        p[1] += s
        p[2] += 1
        ti += i
        ts += s
rows = [("kind", "name", "value", "unit/extra")]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u, data = rr[0], rr[1], rr[2 + which]
keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
for k in keys:
    if k in h:
        i = h.index(k)
        rows.append(("metric", k, data[i], u[i]))
for k in h:
    if "tensor" in k and k not in keys and "pct" in k and ".avg." in k:
        i = h.index(k)
        try:
            nonzero = float(data[i]) != 0.0
        except ValueError:
            nonzero = False
        if nonzero:
            rows.append(("metric", k, data[i], u[i]))
for (f, l), v in sorted(per.items(), key=lambda kv: -kv[1][0])[:30]:
    text = lines[l - 1].strip()[:110] if f == base and l and l <= len(lines) else ""
    rows.append(("line", f"{f}:{l}", f"inst {100 * v[0] / max(ti, 1):.1f}% samples {100 * v[1] / max(ts, 1):.1f}% sass {v[2]}", text))
csv.writer(open(out, "w")).writerows(rows)
print(f"{out}: {len(rows)} rows; instructions {ti}, samples {ts}")
