"""Where the time of one configs[2] pair goes (CUDA events + host wall clock per stage)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfm_registration_b200 as v
from vfm_registration_b200 import api, synth
from scipy.spatial.transform import Rotation as R
import warnings
warnings.simplefilter("ignore")
rng = np.random.default_rng(3)
b, hh, ww, n_map, n_scan = 6, 224, 224, 50_000, 10_000
imgs = torch.from_numpy(rng.integers(1, 255, (b, hh, ww, 3), dtype=np.uint8)).cuda()
kmat = np.array([[200.0, 0, 112.0], [0, 200.0, 112.0], [0, 0, 1.0]])
ts = []
for i in range(b):
    t = np.eye(4); t[:3, :3] = (R.from_euler("z", 60.0 * i, degrees=True) * R.from_euler("yx", [90, -90], degrees=True)).as_matrix().T; ts.append(t)
ks, ts = np.stack([kmat] * b), np.stack(ts)
map_xyz = np.c_[rng.uniform(-20, 20, (n_map, 2)), rng.uniform(-2, 4, n_map)].astype(np.float32)
f = v.ViTFeaturizer("vitl14", seed=4, random_init=True)
map_desc = v.extract_features(imgs, map_xyz, ks, ts, featurizer=f)
rm = v.ResidentMap(torch.from_numpy(map_xyz).cuda(), map_desc)
scan = torch.from_numpy(map_xyz[rng.permutation(n_map)[:n_scan]]).cuda()
kw = dict(min_cos=0.8, mutual=True, ransac_iters=8192, inlier_thresh=0.5, seed=3)
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
tok = f.forward(imgs)
cams = [api.CameraSpec(P=ks[i] @ ts[i][:3], img_hw=(hh, ww), grid_hw=tuple(tok.shape[1:3]), black_mode=1) for i in range(b)]
desc = v.extract_features(imgs, scan, ks, ts, featurizer=f)
print(f"vit forward            {timed(lambda: f.forward(imgs)):.3f} ms")
print(f"project_gather (10k)   {timed(lambda: api.project_gather(scan, cams, [tok[i] for i in range(b)], [imgs[i] for i in range(b)])):.3f} ms")
print(f"extract_features       {timed(lambda: v.extract_features(imgs, scan, ks, ts, featurizer=f)):.3f} ms")
print(f"register_scans         {timed(lambda: v.register_scans(rm, [(scan, desc)], **kw)):.3f} ms")
print(f"whole pair             {timed(lambda: v.register_scans(rm, [(scan, v.extract_features(imgs, scan, ks, ts, featurizer=f))], **kw)):.3f} ms")
ctx = v.get_context(0)
ctx.enable_timing(True)
n = 10
for _ in range(n):
    r = v.register_scans(rm, [(scan, desc)], **kw)[0]
torch.cuda.synchronize()
names = ["forward search", "ransac score", "project", "vit", "reverse search", "normalize", "rerank+exact", "filter_corr", "gather_rows",
         "filter_mutual", "gather_pq+kabsch", "finalize"]
for g, name in enumerate(names):
    ms, c = ctx.group_time_ms(g)
    if c:
        print(f"  {name:18s} {ms / n * 1e3:8.1f} us per pair ({c / n:.1f} launches)")
print("n_corr", len(r.corr), "inliers", r.n_inliers)
