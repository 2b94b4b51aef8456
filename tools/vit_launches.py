#!/usr/bin/env python
"""Per-kernel averages of a ViT forward from an ncu launch list (second half of the launches = steady-state forwards).
    ncu --clock-control none --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,dram__bytes.sum \
        -c 900 --csv --log-file out.csv python tools/one_vit.py vitl14 48;  python tools/vit_launches.py out.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
ki, vi, gi, mi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Metric Name")
recs = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi or not r[0].isdigit():
        continue
    recs.setdefault(int(r[0]), {"name": re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("vfm::", ""), "grid": r[gi]})[r[mi]] = float(r[vi].replace(",", ""))
ids = sorted(recs)
# the last forward: from the last preprocess_kernel on
start = max(i for i in ids if recs[i]["name"].startswith("preprocess"))
seq = [recs[i] for i in ids if i >= start]
per = collections.OrderedDict()
order = 0
for k, r in enumerate(seq):
    # distinguish the GEMMs of a layer by their position after the previous kernel
    prev = seq[k - 1]["name"] if k else ""
    tag = r["name"]
    if "vit_gemm_kernel<2>" in tag:
        tag += " proj" if "attention" in prev else " fc2"
    per.setdefault((tag, r["grid"]), []).append(r)
tot = sum(r["gpu__time_duration.sum"] for r in seq)
print(f"# last forward: {len(seq)} launches, {tot / 1e3:.1f} us (ncu: serialised, caches flushed between launches)")
for (tag, grid), v in per.items():
    t = sum(r["gpu__time_duration.sum"] for r in v) / len(v) / 1e3
    tp = sum(r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) for r in v) / len(v)
    l2 = sum(r.get("lts__t_bytes.sum", 0) for r in v) / len(v) / 1e6
    dr = sum(r.get("dram__bytes.sum", 0) for r in v) / len(v) / 1e6
    print(f"{tag[:34]:34s} {grid:14s} n={len(v):3d} avg {t:8.2f} us  tensor {tp:5.1f} %  L2 {l2:8.1f} MB ({l2 / max(t, 1e-9):6.2f} TB/s)  DRAM {dr:7.1f} MB")
