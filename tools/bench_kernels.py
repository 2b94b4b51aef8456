#!/usr/bin/env python
"""Kernel-level timings (CUDA events on the launching stream, inputs larger than L2 or cycled) used while tuning.
Not the headline bench -- see bench.py.   python tools/bench_kernels.py [match|ransac|project|vit] ..."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfm_registration_b200 as v  # noqa: E402


def time_fn(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_match(shapes=((10000, 50000, 384), (50000, 10000, 384), (4096, 4096, 384), (300, 200000, 384), (20000, 200000, 768))):
    ctx = v.get_context(0)
    for n, m, d in shapes:
        g = torch.Generator(device="cuda").manual_seed(n + m)
        sets = [(torch.randn(n, d, device="cuda", generator=g), torch.randn(m, d, device="cuda", generator=g)) for _ in range(3)]
        for algo in ("tc", "simt") if n * m <= 10000 * 50000 else ("tc",):
            k = [0]

            def run():
                a, b = sets[k[0] % 3]
                k[0] += 1
                v.match_nn(a, b, algo=algo)
            ctx.enable_timing(True)
            ms = time_fn(run, iters=6 if algo == "tc" else 3)
            gms, gl = ctx.group_time_ms(0)
            ctx.enable_timing(False)
            fl = 2.0 * n * m * d
            print(f"match {algo:4s} n={n} m={m} d={d}: call {ms:.3f} ms, GEMM kernel {gms / max(gl, 1):.3f} ms "
                  f"= {fl / (gms / max(gl, 1)) / 1e9:.1f} TFLOP/s", flush=True)


def bench_match_modes(shapes=((10000, 50000, 384), (20000, 200000, 768), (300, 200000, 384))):
    """The candidate-search kernel in its three recording modes against a resident map: top-2 (runner-up needed), top-1,
    top-1 with the caller's cosine gate as the recording floor (what register() with min_cos runs)."""
    from vfm_registration_b200 import synth
    ctx = v.get_context(0)
    for n, m, d in shapes:
        s = synth.make_pair(2, m, n, d)   # 30 % of the queries have a planted match at cosine ~0.9
        rm = v.ResidentMap(torch.from_numpy(s["map_xyz"]).cuda(), torch.from_numpy(s["map_feat"]).cuda())
        a = torch.from_numpy(s["scan_feat"]).cuda()
        fl = 2.0 * n * m * d
        for name, kw in (("top-2", dict(second=True)), ("top-1", dict(second=False)), ("top-1 + gate 0.8", dict(second=False, min_cos=0.8))):
            ctx.enable_timing(True)
            ms = time_fn(lambda: rm.match(a, **kw), iters=10)
            gms, gl = ctx.group_time_ms(0)
            ctx.enable_timing(False)
            print(f"match modes n={n} m={m} d={d} {name:17s}: call {ms:.3f} ms, search kernel {gms / max(gl, 1):.4f} ms "
                  f"= {fl / (gms / max(gl, 1)) / 1e9:.1f} TFLOP/s", flush=True)
        rm.close()


def bench_ransac(shapes=((8192, 3000), (8192, 10000), (65536, 20000), (50000, 300))):
    from vfm_registration_b200 import synth
    ctx = v.get_context(0)
    for h, k in shapes:
        s = synth.make_pair(1, max(k, 4000), k, 8, inlier_frac=0.3)
        rng = np.random.default_rng(0)
        j = np.where(s["perm"] >= 0, s["perm"], rng.integers(0, len(s["map_xyz"]), k))
        corr = torch.from_numpy(np.stack([np.arange(k), j], 1).astype(np.int32)).cuda()
        sx, tx = torch.from_numpy(s["scan_xyz"]).cuda(), torch.from_numpy(s["map_xyz"]).cuda()
        ctx.enable_timing(True)
        ms = time_fn(lambda: v.ransac_kabsch(sx, tx, corr, n_hyp=h, seed=3, thresh=1.0), iters=5)
        gms, gl = ctx.group_time_ms(1)
        ctx.enable_timing(False)
        sc = gms / max(gl, 1)
        print(f"ransac H={h} K={k}: call {ms:.3f} ms, score kernel {sc:.3f} ms = {h * k / sc / 1e6:.1f} G(hyp*corr)/s, "
              f"{h * k * 32 / sc / 1e9:.1f} TFLOP/s fp64-equivalent (16 DFMA-class ops per test), {h / ms / 1e3:.0f} M hyp/s", flush=True)


def bench_project(n=2_000_000, d=384):
    rng = np.random.default_rng(0)
    pts = torch.from_numpy(np.c_[rng.uniform(-20, 20, (n, 2)), rng.uniform(-2, 4, n)].astype(np.float32)).cuda()
    from scipy.spatial.transform import Rotation as R
    cams, toks, imgs = [], [], []
    for i in range(6):
        t = np.eye(4)
        t[:3, :3] = (R.from_euler("z", 60.0 * i, degrees=True) * R.from_euler("yx", [90, -90], degrees=True)).as_matrix().T
        kmat = np.array([[200.0, 0, 112.0], [0, 200.0, 112.0], [0, 0, 1.0]])
        cams.append(v.CameraSpec(P=kmat @ t[:3], img_hw=(224, 224), grid_hw=(16, 16), black_mode=1))
        toks.append(torch.randn(16, 16, d, device="cuda"))
        imgs.append(torch.randint(1, 255, (224, 224, 3), dtype=torch.uint8, device="cuda"))
    ctx = v.get_context(0)
    ctx.enable_timing(True)
    ms = time_fn(lambda: v.project_gather(pts, cams, toks, imgs), iters=10)
    gms, gl = ctx.group_time_ms(2)
    ctx.enable_timing(False)
    k = gms / 13   # device time of one call (3 warm-up + 10 timed calls; a call is one launch, or four on the binned path)
    by = n * (12 + 4 * d + 12) + 6 * 256 * d * 4
    print(f"project_gather N={n} D={d}: call {ms:.3f} ms, kernels {k:.3f} ms ({gl // 13} launches) = {by / k / 1e6:.0f} GB/s algorithmic "
          f"(12 B read + {4 * d} B desc + 12 B index written per point)", flush=True)


def vit_flops(depth, w, t, tp, b):
    """L (24 T W^2 + 4 T^2 W) + 2 T_p 588 W per image."""
    return b * (depth * (24.0 * t * w * w + 4.0 * t * t * w) + 2.0 * tp * 588 * w)


def bench_vit(batches=(6, 48), models=("vits14", "vitb14", "vitl14")):
    from vfm_registration_b200 import features
    ctx = v.get_context(0)
    rng = np.random.default_rng(0)
    for model in models:
        depth, w, heads = features.PRESETS[model]
        f = v.ViTFeaturizer(model, seed=1, random_init=True)
        for b in batches:
            imgs = torch.from_numpy(rng.integers(0, 255, (b, 224, 224, 3), dtype=np.uint8)).cuda()
            ctx.enable_timing(True)
            ms = time_fn(lambda: f.forward(imgs), iters=10)
            gms, gl = ctx.group_time_ms(3)
            ctx.enable_timing(False)
            fl = vit_flops(depth, w, 257, 256, b)
            print(f"vit {model} B={b} (224x224): forward {ms:.3f} ms = {b / ms * 1e3:.0f} img/s, {fl / ms / 1e9:.1f} TFLOP/s "
                  f"({gl // 13} launches per forward)", flush=True)
        del f


def bench_extract(n=10000):
    """BASELINE configs[2]: 6 x (224 x 224) surround images -> ViT -> projection gather for n points."""
    from scipy.spatial.transform import Rotation as R
    rng = np.random.default_rng(0)
    imgs = rng.integers(1, 255, (6, 224, 224, 3), dtype=np.uint8)
    pts = np.c_[rng.uniform(-20, 20, (n, 2)), rng.uniform(-2, 4, n)].astype(np.float32)
    k = np.stack([np.array([[200.0, 0, 112.0], [0, 200.0, 112.0], [0, 0, 1.0]])] * 6)
    ts = []
    for i in range(6):
        t = np.eye(4)
        t[:3, :3] = (R.from_euler("z", 60.0 * i, degrees=True) * R.from_euler("yx", [90, -90], degrees=True)).as_matrix().T
        ts.append(t)
    ts = np.stack(ts)
    for model in ("vits14", "vitl14"):
        f = v.ViTFeaturizer(model, seed=1, random_init=True)
        imgs_d, pts_d = torch.from_numpy(imgs).cuda(), torch.from_numpy(pts).cuda()
        ms = time_fn(lambda: v.extract_features(imgs_d, pts_d, k, ts, featurizer=f), iters=10)
        ms_h = time_fn(lambda: v.extract_features(imgs, pts, k, ts, featurizer=f).cpu(), iters=10)
        print(f"extract_features {model}: 6 images + {n} points: {ms:.3f} ms device-resident, {ms_h:.3f} ms from/to host", flush=True)


def bench_voxel():
    """voxel down-sampling, voxel-map build, 27-voxel nearest neighbour, ICP (NCLT-like sizes)."""
    rng = np.random.default_rng(0)
    for n, cols in ((100_000, 3), (1_000_000, 3), (50_000, 387)):
        pts = torch.from_numpy(np.c_[rng.uniform(-50, 50, (n, 2)), rng.uniform(-2, 8, n), rng.standard_normal((n, cols - 3))]
                               .astype(np.float32)).cuda()
        for vs in (0.5, 1.0):
            ms = time_fn(lambda: v.voxel_down_sample(pts, vs), iters=5)
            kept = v.voxel_down_sample(pts, vs).shape[0]
            print(f"voxel_down_sample N={n} cols={cols} voxel={vs}: {ms:.3f} ms ({kept} kept, {n / ms / 1e3:.0f} M points/s, "
                  f"{(12 * n + 2 * 4 * cols * kept) / ms / 1e6:.0f} GB/s algorithmic)", flush=True)
    map_pts = np.c_[rng.uniform(-50, 50, (200_000, 2)), rng.uniform(-2, 8, 200_000)]
    m = v.VoxelMap(1.0, 20)
    ms = time_fn(lambda: m.build(map_pts), iters=5)
    print(f"VoxelMap.build 200000 points: {ms:.3f} ms incl. the host->device copy ({len(m)} kept)", flush=True)
    scan = map_pts[rng.choice(200_000, 20_000, replace=False)] + rng.normal(0, 0.02, (20_000, 3))
    ang = np.deg2rad(2.0)
    T = np.eye(4)
    T[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]
    T[:3, 3] = [0.3, -0.2, 0.02]
    scan_t = torch.from_numpy((scan - T[:3, 3]) @ T[:3, :3]).cuda()
    ms = time_fn(lambda: m.nearest(scan_t, 3.0), iters=5)
    print(f"VoxelMap.nearest 20000 queries (27 voxels, <= 20 points each): {ms:.3f} ms = {20_000 / ms / 1e3:.1f} M queries/s", flush=True)
    _, info = v.register_frame(scan_t, m, np.eye(4), 6.0, 2.0 / 3.0, return_info=True)
    ms = time_fn(lambda: v.register_frame(scan_t, m, np.eye(4), 6.0, 2.0 / 3.0), iters=5)
    print(f"register_frame 20000 x 200000: {ms:.3f} ms for {info['iterations']} iterations "
          f"({ms / max(info['iterations'], 1) * 1e3:.0f} us per iteration, 2 launches each, host check every 8)", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "match"
    if what == "match":
        bench_match()
    elif what == "modes":
        bench_match_modes()
    elif what == "ransac":
        bench_ransac()
    elif what == "voxel":
        bench_voxel()
    elif what == "project":
        bench_project()
    elif what == "vit":
        bench_vit()
    elif what == "vitl":   # one model, one batch size: python tools/bench_kernels.py vitl 6
        bench_vit(batches=(int(sys.argv[2]) if len(sys.argv) > 2 else 6,), models=("vitl14",))
    elif what == "extract":
        bench_extract()
