#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv [--frames] > profiles/rNN_launches.md
--frames: keep only the launches between the first two L2-flush fills of bench.py (a 256 MB FillFunctor<unsigned char>, one per
timed step): the timed resident-input frame plus the e2e warm-up frames that follow it, i.e. whole frames only; totals are then
reported per frame (frames = k_render_finish launches / passes is ambiguous, so: number of k_occ_box launches / 2 in the 3-pass
scheme, falling back to 1)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
cols, data = rows[hdr], rows[hdr + 1:]
ki, vi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")


def dur_us(r):
    v = float(r[vi].replace(",", ""))
    return v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)


data = [r for r in data if len(r) > vi]
frames = 1.0
if "--frames" in sys.argv:
    fl = [i for i, r in enumerate(data) if "FillFunctor<unsigned char>" in r[ki] and dur_us(r) > 20.0]
    if len(fl) >= 2:
        data = data[fl[0] + 1: fl[1]]
        n_env = sum(1 for r in data if "k_permute_log" in r[ki])
        frames = max(1.0, n_env / 2.0) if n_env else 1.0
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    n = r[ki].split("(")[0][:70]
    agg[n][0] += 1
    agg[n][1] += dur_us(r)
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches / frame | ms / frame | share |\n|---|---:|---:|---:|")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:20]:
    print(f"| `{n}` | {c / frames:.1f} | {t / 1e3 / frames:.3f} | {100 * t / tot:.1f}% |")
print(f"\n{len(data)} launches over {frames:g} frame(s), {tot / 1e3 / frames:.2f} ms of kernel time per frame "
      f"(cold-cache, serialised under ncu: compare shares, not absolutes)")
