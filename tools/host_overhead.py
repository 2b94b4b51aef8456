"""How much host time does one pair cost in vfmreg_register_batch?  Tiny pairs make the GPU work negligible, so the batch
time per pair approximates the enqueue cost (kernel launches, memsets, tensor-map encodes, events)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfm_registration_b200 as v
from vfm_registration_b200 import synth
ctx = v.get_context(0)
for (m, n, d, h) in ((512, 256, 64, 64), (50000, 10000, 384, 8192)):
    s = synth.make_pair(1, m, n, d)
    pair = tuple(torch.from_numpy(s[k]).cuda() for k in ("scan_xyz", "map_xyz", "scan_feat", "map_feat"))
    pairs = [pair] * 64
    for lanes in (1, 3):
        ctx.set_lanes(lanes)
        kw = dict(min_cos=0.8, mutual=True, ransac_iters=h, inlier_thresh=1.0)
        v.register_batch(pairs, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            v.register_batch(pairs, **kw)
        dt = (time.perf_counter() - t0) / (5 * 64)
        print(f"{m}x{n}x{d} H={h} lanes={lanes}: {dt * 1e6:.0f} us per pair (wall, 64 pairs per call)")
