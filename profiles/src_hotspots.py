#!/usr/bin/env python
"""Aggregate warp-stall samples per source line from an ncu report.

usage:  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv ;  python profiles/src_hotspots.py x.csv [top]
Prints the lines holding the most stall samples (with the dominant stall reasons) and a per-file total.
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path)))
    cur_file, hdr = None, None
    per_line = defaultdict(lambda: defaultdict(float))
    text = {}
    line_no = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            i_samp = hdr.index("# Samples")
            stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None:
            continue
        if r[0] != "":
            line_no = r[0]
            text[(cur_file, line_no)] = r[1].strip()
            continue                                    # source line rows carry the aggregate; count SASS rows only
        key = (cur_file, line_no)
        try:
            s = float(r[i_samp] or 0)
        except ValueError:
            continue
        per_line[key]["samples"] += s
        for i, h in stall_cols:
            try:
                per_line[key][h] += float(r[i] or 0)
            except ValueError:
                pass
    total = sum(v["samples"] for v in per_line.values())
    per_file = defaultdict(float)
    for (f, _), v in per_line.items():
        per_file[f] += v["samples"]
    print(f"total samples {total:.0f}")
    for f, s in sorted(per_file.items(), key=lambda kv: -kv[1]):
        print(f"  {f:24s} {s:9.0f}  {100 * s / max(total, 1):5.1f}%")
    print()
    for (f, ln), v in sorted(per_line.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(((h, c) for h, c in v.items() if h != "samples" and c > 0), key=lambda hc: -hc[1])[:3]
        st_s = " ".join(f"{h[6:]}={100 * c / max(v['samples'], 1):.0f}%" for h, c in st)
        print(f"{100 * v['samples'] / max(total, 1):5.1f}%  {f}:{ln:>4s}  {text.get((f, ln), '')[:90]:90s} | {st_s}")


if __name__ == "__main__":
    main()
