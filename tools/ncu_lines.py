#!/usr/bin/env python3
"""Executed warp instructions per CUDA source line (and their opcodes) from `ncu --page source --print-source cuda,sass --csv`.
Usage: ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv > src.csv; tools/ncu_lines.py src.csv [min_millions]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) * 1e6 if len(sys.argv) > 2 else 1e6
cur, ie = None, None
ops, text, seen = collections.defaultdict(collections.Counter), {}, set()
for r in rows:
    if len(r) > 2 and r[0] == "Line No":
        ie = r.index("Instructions Executed")
        continue
    if len(r) < 8 or ie is None:
        continue
    if r[0] != "":
        cur = int(r[0])
        text.setdefault(cur, r[1])
    if r[2] in ("", "...") or r[2] in seen:
        continue
    seen.add(r[2])
    try:
        n = int(r[ie])
    except ValueError:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3])
    ops[cur][m.group(2) if m else r[3]] += n
tot = sum(sum(v.values()) for v in ops.values())
print(f"total {tot / 1e6:.2f} M warp instructions")
for line in sorted(ops):
    s = sum(ops[line].values())
    if s > thr:
        print(f"{line:5d} {s / 1e6:7.2f}  {text[line].strip()[:64]:64s} | " + " ".join(f"{k}:{v / 1e6:.1f}" for k, v in ops[line].most_common(5)))
