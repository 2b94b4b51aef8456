#!/usr/bin/env python
"""In-situ timeline of a ViT forward (CUDA-graph replay, programmatic dependent launches) from per-CTA globaltimer records.
Needs the trace build of the library:
    VFM_BUILD_SUFFIX=trace VFM_NVCC_DEFS=-DVFM_TRACE python -m vfm_registration_b200.build
    VFMREG_LIB=vfm_registration_b200/libvfmreg_b200_trace.so python tools/vit_trace.py vitl14 6 > profiles/r2_vit_trace_b6.txt
Per launch (grouped by kind and by gaps in time): first CTA entry, first CTA past griddepcontrol.wait, last CTA exit."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfm_registration_b200 as v  # noqa: E402
from vfm_registration_b200 import _lib  # noqa: E402

KINDS = {0: "gemm qkv", 1: "gemm fc1+gelu", 2: "gemm partial", 3: "gemm patch", 4: "gemm resid", 10: "layernorm", 11: "attention"}
model = sys.argv[1] if len(sys.argv) > 1 else "vitl14"
b = int(sys.argv[2]) if len(sys.argv) > 2 else 6
lib = _lib.load()
cap = 1 << 18
rec = torch.zeros(cap * 4, dtype=torch.int64, device="cuda")
cur = torch.zeros(1, dtype=torch.int32, device="cuda")
for fn in ("vfmreg_trace_attach_gemm", "vfmreg_trace_attach_ops", "vfmreg_trace_attach_attn"):
    f = getattr(lib, fn)
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
    assert f(rec.data_ptr(), cur.data_ptr(), cap) == 0
feat = v.ViTFeaturizer(model, seed=1, random_init=True)
imgs = torch.randint(0, 255, (b, 224, 224, 3), dtype=torch.uint8, device="cuda")
for _ in range(4):
    feat.forward(imgs)
torch.cuda.synchronize()
cur.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
feat.forward(imgs)
e1.record()
torch.cuda.synchronize()
n = int(cur.item())
r = rec[: 4 * min(n, cap)].cpu().numpy().reshape(-1, 4)
marks = r[(r[:, 0] & 0xFF) >= 100]
r = r[(r[:, 0] & 0xFF) < 100]
kind = r[:, 0] & 0xFF
t0, t1, t2 = r[:, 1], r[:, 2], r[:, 3]
base = t0.min()
# launches: records of one kind whose entry times are closer than any other kind's launch in between -- sort by post-wait
# time and split whenever the kind changes
order = np.argsort(t1, kind="stable")
launches = []
for i in order:
    if launches and launches[-1]["kind"] == kind[i] and t1[i] < launches[-1]["end"] + 200:
        L = launches[-1]
        L["entry"], L["start"], L["end"], L["ctas"] = min(L["entry"], t0[i]), min(L["start"], t1[i]), max(L["end"], t2[i]), L["ctas"] + 1
    else:
        launches.append({"kind": int(kind[i]), "entry": t0[i], "start": t1[i], "end": t2[i], "ctas": 1, "cta_us": 0.0, "last_start": t1[i]})
    launches[-1]["cta_us"] += (t2[i] - t1[i]) / 1e3
    launches[-1]["last_start"] = max(launches[-1]["last_start"], t1[i])
print(f"# {model} B={b}: forward {e0.elapsed_time(e1) * 1e3:.1f} us by CUDA events (trace build), {n} CTA records, {len(launches)} launches seen")
print(f"# {'kernel':16s} {'CTAs':>5s} {'entry':>9s} {'start':>9s} {'end':>9s} {'busy':>7s} {'gap':>6s}   (us; entry = first CTA running, start = past griddepcontrol.wait, gap = start - previous end)")
prev_end = None
per = {}
for L in launches[: 7 * 3 + 3] + [None] + launches[-9:]:
    if L is None:
        print("  ...")
        continue
    gap = (L["start"] - prev_end) / 1e3 if prev_end is not None else 0.0
    print(f"  {KINDS.get(L['kind'], str(L['kind'])):16s} {L['ctas']:5d} {(L['entry'] - base) / 1e3:9.2f} {(L['start'] - base) / 1e3:9.2f} "
          f"{(L['end'] - base) / 1e3:9.2f} {(L['end'] - L['start']) / 1e3:7.2f} {gap:6.2f}")
    prev_end = L["end"]
prev_end = None
seq = {}
for k, L in enumerate(launches):
    name = KINDS.get(L["kind"], str(L["kind"]))
    if L["kind"] in (2, 4):
        name += " (proj)" if k and launches[k - 1]["kind"] == 11 else " (fc2)"
    busy = (L["end"] - L["start"]) / 1e3
    gap = (L["start"] - prev_end) / 1e3 if prev_end is not None else 0.0
    s = seq.setdefault(name, [0, 0.0, 0.0, 0.0, 0.0])
    s[0] += 1
    s[1] += busy
    s[2] += gap
    s[3] += L["cta_us"] / L["ctas"]
    s[4] += (L["last_start"] - L["start"]) / 1e3
    prev_end = L["end"]
print("# per kernel type: launches, mean busy us (first start -> last exit), mean gap before it")
tot = 0.0
for name, (c, bs, gp, cu, ls) in seq.items():
    print(f"  {name:22s} n={c:3d} busy {bs / c:7.2f}  gap {gp / c:6.2f}   total {bs + gp:8.1f}   mean CTA time {cu / c:6.2f}  last CTA starts +{ls / c:5.2f}")
    tot += bs + gp
if len(marks):   # trace_mark() timestamps of CTA 0 (kernel-internal phases), first 40, relative to the first
    m = marks[np.argsort(marks[:, 1])]
    first = [x for x in m if x[0] == 100]
    if first:
        sel = m[(m[:, 1] >= first[0][1])][:40]
        print("# marks of CTA 0 (kind: us since the first):", " ".join(f"{int(k)}:{(tt - sel[0][1]) / 1e3:.2f}" for k, tt in zip(sel[:, 0], sel[:, 1])))
print(f"# sum {tot:.1f} us (final norm / preprocess / cls rows are not traced)")
dist = os.environ.get("VIT_TRACE_DIST")   # kernel kind (e.g. 11 = attention): per-CTA start / duration spread of its first launches
if dist:
    shown = 0
    for L in launches:
        if L["kind"] != int(dist):
            continue
        sel = (kind == L["kind"]) & (t1 >= L["start"]) & (t2 <= L["end"])
        st, du = (t1[sel] - L["start"]) / 1e3, (t2[sel] - t1[sel]) / 1e3
        en = (t2[sel] - L["start"]) / 1e3
        q = lambda a: " ".join(f"{x:.1f}" for x in np.percentile(a, [0, 10, 50, 90, 100]))
        print(f"# kind {dist} launch {shown}: {sel.sum()} CTAs; start after first [{q(st)}] us, duration [{q(du)}], exit [{q(en)}] (min p10 p50 p90 max)")
        shown += 1
        if shown == 4:
            break

