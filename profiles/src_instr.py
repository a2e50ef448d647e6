#!/usr/bin/env python
"""Warp-level instruction counts per source line from an ncu source-page CSV (cuda,sass view).
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv ; python profiles/src_instr.py x.csv [top]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur, hdr = None, None
per = defaultdict(float)
text = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; ii = hdr.index("Instructions Executed"); continue
    if hdr is None or not r[0].isdigit():
        continue
    try:
        v = float(r[ii])
    except Exception:
        continue
    per[(cur, int(r[0]))] += v
    text[(cur, int(r[0]))] = r[1].strip()[:110]
tot = sum(per.values())
print(f"total warp instructions {tot:.0f}")
for (f, l), v in sorted(per.items(), key=lambda x: -x[1])[:top]:
    print(f"{100 * v / tot:5.1f}%  {f}:{l:4d}  {text[(f, l)]}")
