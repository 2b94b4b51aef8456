#!/usr/bin/env python
"""Timeline of the candidate-search kernels inside register_batch (trace build, per-CTA globaltimer records):
    VFM_BUILD_SUFFIX=trace VFM_NVCC_DEFS=-DVFM_TRACE python -m vfm_registration_b200.build
    VFMREG_LIB=vfm_registration_b200/libvfmreg_b200_trace.so python tools/match_trace.py
Prints, per search launch, start / end and the gap since the previous search ended: how busy the search stream is."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfm_registration_b200 as v  # noqa: E402
from vfm_registration_b200 import _lib, synth  # noqa: E402

lib = _lib.load()
cap = 1 << 18
rec = torch.zeros(cap * 4, dtype=torch.int64, device="cuda")
cur = torch.zeros(1, dtype=torch.int32, device="cuda")
f = lib.vfmreg_trace_attach_match
f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
assert f(rec.data_ptr(), cur.data_ptr(), cap) == 0
dev = torch.device("cuda", 0)
pairs = []
for k in range(4):
    sc = synth.make_scene(1000 + k, 50_000, 5, 10_000, 384)
    mx, mf = torch.from_numpy(sc["map_xyz"]).to(dev), torch.from_numpy(sc["map_feat"]).to(dev)
    pairs += [(torch.from_numpy(s["scan_xyz"]).to(dev), mx, torch.from_numpy(s["scan_feat"]).to(dev), mf) for s in sc["scans"]]
kw = dict(min_cos=0.8, mutual=True, ransac_iters=8192, inlier_thresh=1.0, seed=42)
for _ in range(3):
    v.register_batch(pairs, **kw)
torch.cuda.synchronize()
cur.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
v.register_batch(pairs, **kw)
e1.record()
torch.cuda.synchronize()
n = int(cur.item())
r = rec[: 4 * n].cpu().numpy().reshape(-1, 4)
kind, t1, t2 = r[:, 0] & 0xFF, r[:, 2], r[:, 3]
order = np.argsort(t1, kind="stable")
launches = []
for i in order:
    if launches and launches[-1][0] == kind[i] and t1[i] < launches[-1][2]:
        L = launches[-1]
        L[1], L[2] = min(L[1], t1[i]), max(L[2], t2[i])
    else:
        launches.append([int(kind[i]), t1[i], t2[i]])
base = launches[0][1]
print(f"# one step of 20 pairs: {e0.elapsed_time(e1) * 1e3:.0f} us (trace build); {len(launches)} search launches")
busy = 0.0
prev = None
for k, s, e in launches:
    gap = (s - prev) / 1e3 if prev is not None else 0.0
    busy += (e - s) / 1e3
    print(f"  {'F' if k == 20 else 'P'}  start {(s - base) / 1e3:8.1f}  end {(e - base) / 1e3:8.1f}  busy {(e - s) / 1e3:6.1f}  gap {gap:6.1f}")
    prev = e
span = (launches[-1][2] - base) / 1e3
print(f"# search stream: busy {busy:.0f} us of {span:.0f} us ({100 * busy / span:.0f} %)")
