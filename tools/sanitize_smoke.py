import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
import vfm_registration_b200 as v
from vfm_registration_b200 import synth
ONLY_VIT = len(sys.argv) > 1 and sys.argv[1] == "vit"   # `sanitize_smoke.py vit`: the ViT section alone


def vit_section():
    f = v.ViTFeaturizer("vits14", seed=1, random_init=True)
    imgs = torch.randint(0, 255, (2, 70, 82, 3), dtype=torch.uint8, device="cuda")
    for _ in range(3):   # eager, captured, replayed
        tok = f.forward(imgs)
    print("vit", tuple(tok.shape), float(tok.abs().mean()))
    # K kept whole in proj / fc2 -> the fused residual epilogue (EPI_F32_RESID) and the LayerNorm without a pending residual;
    # 129 tokens per image -> one full query block + a one-row last block (the CUDA-core path of the attention kernel)
    os.environ["VFMREG_VIT_PLAN"] = "proj:96:1,fc2:96:1"
    f2 = v.ViTFeaturizer("vits14", seed=1, random_init=True)
    imgs2 = torch.randint(0, 255, (3, 224, 112, 3), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        tok2 = f2.forward(imgs2)
    del os.environ["VFMREG_VIT_PLAN"]
    print("vit resid", tuple(tok2.shape), float(tok2.abs().mean()))


if ONLY_VIT:
    vit_section()
    sys.exit(0)
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
# round 2: resident maps / scenes, the Open3D-style score, top-n selection, and the ViT forward (tcgen05 GEMMs + attention)
rm = v.ResidentMap(dev[1], dev[3])
rr = v.register_scans(rm, [(dev[0], dev[2]), (dev[0], dev[2])], min_cos=0.8, mutual=True, ransac_iters=512, inlier_thresh=1.0)
print("scans", [x.n_inliers for x in rr])
rm.close()
sel = v.select_smallest(m, 500)
print("select", tuple(sel.shape))
vit_section()
