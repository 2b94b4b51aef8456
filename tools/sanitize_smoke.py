import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
import vfm_registration_b200 as v
from vfm_registration_b200 import synth
s = synth.make_pair(3, 3000, 1500, 384)
r = v.register(s["scan_xyz"], s["map_xyz"], s["scan_feat"], s["map_feat"], min_cos=0.8, mutual=True, ransac_iters=1024, inlier_thresh=1.0)
print("register", r.n_inliers, len(r.corr))
dev = [torch.from_numpy(s[k]).cuda() for k in ("scan_xyz", "map_xyz", "scan_feat", "map_feat")]
rs = v.register_batch([dev, dev, dev, dev], min_cos=0.8, mutual=True, ransac_iters=512, inlier_thresh=1.0)
print("batch", [x.n_inliers for x in rs])
m = v.match_nn(dev[2], dev[3], mutual=True)
print("match", int(m.idx01.sum()))
pts = np.random.default_rng(0).uniform(-20, 20, (20000, 3))
ds = v.voxel_down_sample(pts, 1.0)
vm = v.VoxelMap(1.0, 20); vm.build(pts)
T = v.register_frame(pts[:3000] + 0.05, vm, np.eye(4), 3.0, 0.6, max_iterations=5)
T2 = v.register_frame_vfm(pts[:3000] + 0.05, vm, pts[:200] + 0.05, pts[:200], np.eye(4), 3.0, 0.6, max_iterations=6)
print("voxel", ds.shape, len(vm), T[0, 3], T2[0, 3])
