#!/usr/bin/env python3
"""Summarise an .ncu-rep of the cost kernel: headline metrics, dynamic SASS mix, hottest basic blocks."""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for i, h in enumerate(hdr):
    if h in want or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        try:
            if "issue_stalled" in h and float(vals[i]) < 0.08:
                continue
        except ValueError:
            pass
        print(f"{h:90s} {units[i]:12s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
i_src, i_ex = hdr.index("Source"), hdr.index("Instructions Executed")
ops, tot = collections.Counter(), 0
for r in rows[2:]:
    if len(r) <= i_ex:
        continue
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[i_src])
    if not m:
        continue
    op, n = m.group(1), int(r[i_ex])
    key = op.split(".")[0] if not op.startswith("IMAD.") else ".".join(op.split(".")[:2])
    ops[key] += n
    tot += n
print("dynamic warp instructions:", tot)
print(", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in ops.most_common(18)))
