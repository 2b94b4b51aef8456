#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised: compare SHARES).
    python tools/launch_summary.py gpurun_out/r2c_launches.csv [skip_first_n_launches] > profiles/r2_launches_summary.txt"""
import collections
import csv
import re
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path, errors="replace")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
per = collections.OrderedDict()
n = 0
for r in rows[hdr + 1:]:
    if len(r) <= vi or not r[0].isdigit():
        continue
    n += 1
    if n <= skip:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("vfm::", "")
    key = (name, r[gi], r[bi])
    per.setdefault(key, []).append(float(r[vi].replace(",", "")) / 1e3)
tot = sum(sum(v) for v in per.values())
print(f"# {path}: {n - skip} launches, {tot:.1f} us in total (ncu, serialised, cold cache)")
print(f"{'kernel':58s} {'grid':>14s} {'block':>12s} {'n':>5s} {'avg us':>9s} {'sum us':>10s} {'share':>6s}")
for (name, g, b), v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
    print(f"{name[:58]:58s} {g:>14s} {b:>12s} {len(v):5d} {sum(v) / len(v):9.2f} {sum(v):10.1f} {100 * sum(v) / tot:5.1f}%")
