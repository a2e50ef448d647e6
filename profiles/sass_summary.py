#!/usr/bin/env python
"""SASS opcode summary of libenvidr_b200.so (no GPU needed): python profiles/sass_summary.py > profiles/r02_sass.md
Counts, per kernel, the Blackwell tensor-core / TMA / mbarrier mnemonics of B200_PROFILING.md: UTCHMMA = tcgen05.mma kind::f16,
UTCQMMA = kind::f8f6f4, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "envidr_b200", "_lib", "libenvidr_b200.so")], capture_output=True, text=True).stdout
rows = []
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    cnt, n = collections.Counter(), 0
    for l in f.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if not m:
            continue
        n += 1
        op = m.group(1)
        cnt[op.split(".")[0]] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            cnt["UTCHMMA.2CTA"] += 1
    rows.append((subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0], n, cnt))
cols = ["UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "SYNCS", "F2FP", "FFMA"]
print("| kernel | SASS instr | " + " | ".join(cols) + " |\n|---|---:|" + "---:|" * len(cols))
for name, n, c in sorted(rows, key=lambda r: -r[1]):
    if c["UTCHMMA"] + c["UTCQMMA"] + c["UBLKCP"] + c["LDTM"] == 0:
        continue
    print(f"| `{name[:60]}` | {n} | " + " | ".join(str(c[k]) for k in cols) + " |")
print(f"| **whole library ({len(rows)} kernels)** | {sum(r[1] for r in rows)} | " + " | ".join(str(sum(r[2][k] for r in rows)) for k in cols) + " |")
