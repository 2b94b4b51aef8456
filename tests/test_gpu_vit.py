"""GPU numerics of the DINOv2 forward (bf16 tensor-core GEMMs, fp32 accumulation / residual stream) against the
torch-CPU float32 oracle with the same seeded random weights (oracle/vit.py, itself checked against
transformers.Dinov2Model).  Tolerances are for channel-normalised (unit-variance) features and are those of bf16
operands through `depth` layers; they are written next to each assertion."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import vit as ovit  # noqa: E402


@pytest.fixture(scope="module")
def vfm():
    assert torch.cuda.is_available()
    import vfm_registration_b200 as v
    v.get_context(0)
    return v


def _images(rng, b, h, w):
    img = rng.integers(0, 255, (b, h, w, 3), dtype=np.uint8)
    # smooth structure so that the resize matters
    yy, xx = np.mgrid[0:h, 0:w]
    img[..., 0] = (128 + 100 * np.sin(xx / 7.0 + yy / 11.0)).astype(np.uint8)
    img[:, : h // 5, : w // 4] = 0
    return img


@pytest.mark.parametrize("model,depth,hw,b", [("vits14", 12, (224, 224), 2), ("vits14", 12, (70, 82), 1), ("vitb14", 12, (224, 224), 1),
                                              ("vitl14", 24, (224, 224), 1),
                                              ("vits14", 12, (224, 112), 3),   # 16 x 8 patches + CLS = 129 tokens: one full query block + 1 row
                                              ("vits14", 12, (224, 320), 2)])   # 16 x 22 patches + CLS = 353 tokens: the mma.sync attention (> 320 tokens)
def test_vit_forward_vs_oracle(vfm, model, depth, hw, b):
    cfg = ovit.CONFIGS[model]
    sd = ovit.make_weights(cfg, seed=11)
    rng = np.random.default_rng(5)
    imgs = _images(rng, b, *hw)
    f = vfm.ViTFeaturizer(model, sd)
    got = f.forward(imgs).cpu()
    x = torch.stack([ovit.preprocess(im) for im in imgs])
    want = ovit.forward(sd, cfg, x)
    assert got.shape == want.shape == (b, 16, ovit.patch_grid(*hw)[1], cfg.width)
    err = (got - want).abs()
    cos = torch.nn.functional.cosine_similarity(got.flatten(0, 2), want.flatten(0, 2), dim=1)
    # bf16 operands (2^-9 relative) through `depth` residual blocks on unit-variance outputs
    assert cos.min() > 0.999, float(cos.min())
    assert err.mean() < 0.02 and err.max() < 0.25, (float(err.mean()), float(err.max()))


def test_many_images_fused_residual_equals_partial_fold(vfm, monkeypatch):
    """With many images the proj / fc2 GEMMs keep K whole and add into the residual stream in their epilogue
    (EPI_F32_RESID); with few, K is split and the next LayerNorm folds the partial sums.  Same expression either way:
    the two forwards agree bit for bit, and both match the oracle."""
    cfg = ovit.CONFIGS["vits14"]
    sd = ovit.make_weights(cfg, seed=3)
    imgs = _images(np.random.default_rng(8), 64, 224, 224)
    imgs[1:] = np.roll(imgs[1:], 17, axis=2) ^ np.arange(63, dtype=np.uint8)[:, None, None, None]
    got = vfm.ViTFeaturizer("vits14", sd).forward(imgs).cpu()
    monkeypatch.setenv("VFMREG_VIT_RESID", "0")
    folded = vfm.ViTFeaturizer("vits14", sd).forward(imgs).cpu()
    assert torch.equal(got, folded)
    pick = [0, 31, 63]
    want = ovit.forward(sd, cfg, torch.stack([ovit.preprocess(imgs[i]) for i in pick]))
    cos = torch.nn.functional.cosine_similarity(got[pick].flatten(0, 2), want.flatten(0, 2), dim=1)
    assert cos.min() > 0.999, float(cos.min())   # bf16 operands through 12 blocks, as in test_vit_forward_vs_oracle


def test_image_feature_generator_compat(vfm):
    gen = vfm.ImageFeatureGenerator("dinov2", use_featup=False, seed=1, random_init=True)
    img = _images(np.random.default_rng(0), 1, 70, 82)[0]
    f = gen.get_image_features(img)
    assert f.shape == (16, 18, 384) and f.dtype == np.float32          # (16, patch_w, C) like the reference
    assert abs(float(f.mean())) < 0.05 and abs(float(f.std()) - 1.0) < 0.05   # identity-affine ChannelNorm output
    up = gen.get_image_features(img, upsample=True)
    assert up.shape == (70, 82, 384)
    with pytest.raises(ValueError, match="Unsupported foundation model"):
        vfm.ImageFeatureGenerator("maskclip")


def test_extract_features_end_to_end(vfm):
    """BASELINE configs[2] shape: 6 surround images -> ViT -> projection gather; checked against the oracle chain
    (oracle ViT features -> bilinear upsample -> create_descriptors)."""
    from oracle import project
    from scipy.spatial.transform import Rotation as R
    rng = np.random.default_rng(3)
    b, h, w, n = 6, 224, 224, 4000
    imgs = _images(rng, b, h, w)
    pts = np.c_[rng.uniform(-20, 20, (n, 2)), rng.uniform(-2, 4, n)].astype(np.float32)
    kmat = np.array([[200.0, 0, 112.0], [0, 200.0, 112.0], [0, 0, 1.0]])
    ks = np.stack([kmat] * b)
    ts = []
    for i in range(b):
        t = np.eye(4)
        t[:3, :3] = (R.from_euler("z", 60.0 * i, degrees=True) * R.from_euler("yx", [90, -90], degrees=True)).as_matrix().T
        ts.append(t)
    ts = np.stack(ts)
    cfg = ovit.CONFIGS["vits14"]
    sd = ovit.make_weights(cfg, seed=2)
    f = vfm.ViTFeaturizer("vits14", sd)
    desc = vfm.extract_features(imgs, pts, ks, ts, featurizer=f).cpu().numpy()
    x = torch.stack([ovit.preprocess(im) for im in imgs])
    tok = ovit.forward(sd, cfg, x).numpy()
    feats = {i: np.ascontiguousarray(project.upsample_bilinear(np.ascontiguousarray(tok[i].transpose(2, 0, 1)), h, w).transpose(1, 2, 0))
             for i in range(b)}
    images = {i: imgs[i] for i in range(b)}
    fn = lambda pcl_h, image, cam: project.project_pinhole(pcl_h[:3].T, ks[cam], ts[cam], h, w, image=image)  # noqa: E731
    want = project.create_descriptors(images, feats, fn, pts)
    seen = np.abs(want).sum(1) > 0
    assert 0.3 < seen.mean() <= 1.0
    assert np.array_equal(np.abs(desc).sum(1) > 0, seen)                 # identical visible set
    cos = (desc[seen] * want[seen]).sum(1) / (np.linalg.norm(desc[seen], axis=1) * np.linalg.norm(want[seen], axis=1))
    assert cos.min() > 0.999 and np.abs(desc - want).max() < 0.25        # bf16 ViT tolerance, as above


def test_config3_images_to_transform(vfm):
    """BASELINE configs[2] plumbing end to end: 6 surround images -> ViT -> projection gather for a 'map' cloud and for a
    'scan' (a rigidly moved, noisy subset of it seen by the same rig) -> match -> RANSAC recovers the planted SE(3)."""
    from scipy.spatial.transform import Rotation as R
    rng = np.random.default_rng(12)
    b, h, w = 6, 224, 224
    imgs = rng.integers(1, 255, (b, h, w, 3), dtype=np.uint8)
    kmat = np.array([[200.0, 0, 112.0], [0, 200.0, 112.0], [0, 0, 1.0]])
    ts = []
    for i in range(b):
        t = np.eye(4)
        t[:3, :3] = (R.from_euler("z", 60.0 * i, degrees=True) * R.from_euler("yx", [90, -90], degrees=True)).as_matrix().T
        ts.append(t)
    ks, ts = np.stack([kmat] * b), np.stack(ts)
    n_map, n_scan = 6000, 2000
    map_xyz = np.c_[rng.uniform(-20, 20, (n_map, 2)), rng.uniform(-2, 4, n_map)].astype(np.float32)
    f = vfm.ViTFeaturizer("vits14", seed=4, random_init=True)
    map_desc = vfm.extract_features(imgs, map_xyz, ks, ts, featurizer=f)
    sel = rng.permutation(n_map)[:n_scan]
    t_gt = np.eye(4)
    t_gt[:3, :3] = R.from_euler("zyx", [25.0, 1.0, -2.0], degrees=True).as_matrix()
    t_gt[:3, 3] = [4.0, -3.0, 0.5]
    t_inv = np.linalg.inv(t_gt)
    # the scan is the same physical points expressed in the moved sensor frame; its descriptors come from the same pixels
    scan_xyz = (map_xyz[sel].astype(np.float64) @ t_inv[:3, :3].T + t_inv[:3, 3] + rng.normal(0, 0.01, (n_scan, 3))).astype(np.float32)
    scan_desc = map_desc[torch.from_numpy(sel).cuda()]
    seen = (scan_desc.abs().sum(1) > 0).cpu().numpy()
    assert seen.mean() > 0.3
    r = vfm.register(torch.from_numpy(scan_xyz).cuda(), torch.from_numpy(map_xyz).cuda(), scan_desc, map_desc, min_cos=0.8,
                     ransac_iters=4096, inlier_thresh=0.5, seed=3)
    # unseen points carry zero descriptors: they can never pass the 0.8 cosine gate (VoxelHashMap.cpp:503)
    assert set(r.corr[:, 0].tolist()) <= set(np.nonzero(seen)[0].tolist())
    rte = np.linalg.norm(r.T[:3, 3] - t_gt[:3, 3])
    rre = np.degrees(np.arccos(np.clip((np.trace(r.T[:3, :3].T @ t_gt[:3, :3]) - 1) / 2, -1, 1)))
    assert rte < 0.2 and rre < 0.5, (rte, rre)


def test_vit_graph_replay_matches_eager(vfm):
    """The second and later calls with a given shape replay a captured CUDA graph: results must be bit-identical to the
    eager first call, for changing inputs, interleaved shapes and a batch that forces the staging buffers to grow."""
    rng = np.random.default_rng(2)
    f = vfm.ViTFeaturizer("vits14", seed=5, random_init=True)
    a = torch.from_numpy(_images(rng, 2, 224, 224)).cuda()
    b = torch.from_numpy(_images(rng, 2, 224, 224)).cuda()
    c = torch.from_numpy(_images(rng, 1, 70, 82)).cuda()
    ea, ec = f.forward(a).clone(), f.forward(c).clone()       # eager
    ga1 = f.forward(a).clone()                                 # captures + replays
    gb = f.forward(b).clone()                                  # replay, other input
    gc = f.forward(c).clone()                                  # second shape: captures
    ga2 = f.forward(a).clone()
    assert torch.equal(ea, ga1) and torch.equal(ea, ga2) and torch.equal(ec, gc)
    assert not torch.equal(ga1, gb)
    big = torch.from_numpy(_images(rng, 5, 224, 224)).cuda()
    e_big = f.forward(big).clone()
    g_big = f.forward(big).clone()
    assert torch.equal(e_big, g_big) and torch.equal(f.forward(a), ea)


def test_nn_indices_survive_the_bf16_forward(vfm):
    """What the registration consumes is not the tokens but the nearest-neighbour indices computed from them.  Descriptors
    gathered from the GPU forward (bf16 GEMM operands) and from the fp32 oracle forward of the same seeded ViT-S/14 are
    matched (scan vs map, cosine gate 0.8 as VoxelHashMap.cpp:486-511): at least 99 % of the gated queries pick the same map
    row, and the two gated sets agree to within 1 %."""
    cfg = ovit.CONFIGS["vits14"]
    sd = ovit.make_weights(cfg, seed=21)
    rng = np.random.default_rng(9)
    imgs = _images(rng, 3, 224, 224)
    f = vfm.ViTFeaturizer("vits14", sd)
    tok_gpu = f.forward(imgs).cpu().numpy()                                     # (3, 16, 16, 384)
    tok_ref = ovit.forward(sd, cfg, torch.stack([ovit.preprocess(im) for im in imgs])).numpy()

    def sample(tok, cam, gy, gx):   # bilinear sample of the token grid at continuous grid coordinates
        y0, x0 = np.floor(gy).astype(int), np.floor(gx).astype(int)
        y1, x1 = np.minimum(y0 + 1, 15), np.minimum(x0 + 1, 15)
        wy, wx = (gy - y0)[:, None], (gx - x0)[:, None]
        t = tok[cam]
        return ((1 - wy) * ((1 - wx) * t[y0, x0] + wx * t[y0, x1]) + wy * ((1 - wx) * t[y1, x0] + wx * t[y1, x1])).astype(np.float32)

    m, n = 6000, 2000
    cam_m, gy_m, gx_m = rng.integers(0, 3, m), rng.uniform(0, 15, m), rng.uniform(0, 15, m)
    pick = rng.choice(m, n, replace=False)           # every scan point sits near one map point (a few hundredths of a cell)
    cam_s, gy_s, gx_s = cam_m[pick], np.clip(gy_m[pick] + rng.normal(0, 0.03, n), 0, 15), np.clip(gx_m[pick] + rng.normal(0, 0.03, n), 0, 15)
    res = {}
    for name, tok in (("gpu", tok_gpu), ("ref", tok_ref)):
        mp = np.concatenate([sample(tok, c, gy_m[cam_m == c], gx_m[cam_m == c]) for c in range(3)])
        order = np.concatenate([np.nonzero(cam_m == c)[0] for c in range(3)])
        map_feat = np.empty_like(mp)
        map_feat[order] = mp
        sc = np.empty((n, mp.shape[1]), dtype=np.float32)
        for c in range(3):
            sc[cam_s == c] = sample(tok, c, gy_s[cam_s == c], gx_s[cam_s == c])
        r = vfm.match_nn(sc, map_feat)
        res[name] = (r.idx01.cpu().numpy(), r.sim01.cpu().numpy())
    gated_gpu, gated_ref = res["gpu"][1] >= 0.8, res["ref"][1] >= 0.8
    both = gated_gpu & gated_ref
    assert both.sum() > 0.9 * n, int(both.sum())                                   # the planted neighbours clear the gate
    assert (gated_gpu != gated_ref).mean() < 0.01
    agree = (res["gpu"][0][both] == res["ref"][0][both]).mean()
    assert agree >= 0.99, float(agree)
