"""One workload for ncu: python tools/one_register.py configs1|configs3|refshape [calls]
configs1 = one scene (1 map + 5 scans) through register_batch, configs3 = one 200k x 20k x 768 ratio-test pair,
refshape = 300 queries against a resident 200k-point map."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfm_registration_b200 as v  # noqa: E402
from vfm_registration_b200 import synth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "configs1"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
if what == "configs1":
    pairs = []
    for k in range(int(os.environ.get("SCENES", "1"))):
        sc = synth.make_scene(1000 + k, 50_000, 5, 10_000, 384)
        mx, mf = torch.from_numpy(sc["map_xyz"]).to(dev), torch.from_numpy(sc["map_feat"]).to(dev)
        pairs += [(torch.from_numpy(s["scan_xyz"]).to(dev), mx, torch.from_numpy(s["scan_feat"]).to(dev), mf) for s in sc["scans"]]
    if os.environ.get("ONE_LANE"):
        v.get_context(0).set_lanes(1)
    for _ in range(calls):
        v.register_batch(pairs, min_cos=0.8, mutual=True, ransac_iters=8192, inlier_thresh=1.0, seed=42)
elif what == "configs3":
    s = synth.make_pair(4, 200_000, 20_000, 768, sigma_f=0.02)
    args = tuple(torch.from_numpy(s[k]).to(dev) for k in ("scan_xyz", "map_xyz", "scan_feat", "map_feat"))
    for _ in range(calls):
        v.register(*args, min_cos=None, ratio=0.9, ransac_iters=65536, inlier_thresh=1.0, seed=4)
else:
    rng = np.random.default_rng(77)
    b = torch.from_numpy(rng.standard_normal((200_000, 384)).astype(np.float32)).to(dev)
    a = torch.from_numpy(rng.standard_normal((300, 384)).astype(np.float32)).to(dev)
    rm = v.ResidentMap(torch.zeros(200_000, 3, device=dev), b)
    for _ in range(calls):
        rm.match(a, min_cos=0.8, second=False)
torch.cuda.synchronize()
