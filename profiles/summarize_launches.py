#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
cols, data = rows[hdr], rows[hdr + 1:]
ki, vi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    n = r[ki].split("(")[0][:70]
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:20]:
    print(f"| `{n}` | {c} | {t / 1e3:.3f} | {100 * t / tot:.1f}% |")
print(f"\n{len(data)} launches, {tot / 1e3:.2f} ms total (cold-cache, serialised under ncu: compare shares, not absolutes)")
