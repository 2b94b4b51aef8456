"""Per-stage device time of a pair inside a batch (CUDA events on the stream each stage is launched on; the interval runs
from the moment the stage reaches the head of its stream to its completion, so it includes waiting for SM resources beside
the other lanes).  python tools/stage_times.py [lanes] [scenes]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfm_registration_b200 as v  # noqa: E402
from vfm_registration_b200 import synth  # noqa: E402

lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 5
scenes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda", 0)
ctx = v.get_context(0)
ctx.set_lanes(lanes)
pairs = []
for k in range(scenes):
    sc = synth.make_scene(1000 + k, 50_000, 5, 10_000, 384)
    mx, mf = torch.from_numpy(sc["map_xyz"]).to(dev), torch.from_numpy(sc["map_feat"]).to(dev)
    pairs += [(torch.from_numpy(s["scan_xyz"]).to(dev), mx, torch.from_numpy(s["scan_feat"]).to(dev), mf) for s in sc["scans"]]
kw = dict(min_cos=0.8, mutual=True, ransac_iters=8192, inlier_thresh=1.0, seed=42)
for _ in range(3):
    v.register_batch(pairs, **kw)
torch.cuda.synchronize()
ctx.enable_timing(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
steps = 5
for _ in range(steps):
    v.register_batch(pairs, **kw)
e1.record()
torch.cuda.synchronize()
names = ["forward search", "ransac score", "project", "vit", "reverse search", "normalize", "rerank+exact", "filter_corr", "gather_rows",
         "filter_mutual", "gather_pq+kabsch", "finalize"]
n_pairs = steps * len(pairs)
print(f"lanes {lanes}: {e0.elapsed_time(e1) / n_pairs * 1e3:.1f} us per pair ({n_pairs / e0.elapsed_time(e1) * 1e3:.0f} pairs/s)")
tot = 0.0
for g, name in enumerate(names):
    ms, n = ctx.group_time_ms(g)
    if n:
        print(f"  {name:18s} {ms / n_pairs * 1e3:8.1f} us per pair  ({n / n_pairs:.2f} launches per pair)")
        tot += ms / n_pairs * 1e3
print(f"  sum {tot:.1f} us per pair")
