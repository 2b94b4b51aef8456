"""Short measurements of the BASELINE.json configurations that are not the headline (bench.py `extras`, N = 1):

  configs[0]  4096 x 4096 points, 384-d, 1024 hypotheses, cosine gate only -- the reference's own CPU-runnable case
  configs[2]  6 x (224 x 224) images -> DINOv2 ViT-L/14 (tcgen05) -> projection gather -> match -> RANSAC
  configs[3]  200k map x 20k scan, 768-d, 65536 hypotheses, ratio test 0.9
  reference_shape  ~300 voxelised queries against a resident 200k-point map: the HBM-bound search shape

Every leg is timed on the device with CUDA events per iteration; the L2 is flushed (a 256 MB buffer is rewritten) between
iterations whose inputs are smaller than the L2.  Also runnable alone:  python -m tools.bench_extras [name ...]"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_flush = None


def _time(fn, iters, warmup=3, flush=True):
    """ms per call (mean over iters), CUDA events per iteration on the current stream; optional L2 flush before each one."""
    global _flush
    if flush and _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        if flush:
            _flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    return total / iters


def _pin(a):
    return torch.from_numpy(a).pin_memory().numpy()


def config0(v, ctx, dev, peaks):
    from oracle import cref, match
    from vfm_registration_b200 import synth
    kw = dict(min_cos=0.8, ransac_iters=1024, inlier_thresh=1.0, seed=1)
    pairs = [synth.make_pair(1 + i, 4096, 4096, 384) for i in range(3)]
    keys = ("scan_xyz", "map_xyz", "scan_feat", "map_feat")
    dpairs = [tuple(torch.from_numpy(s[k]).to(dev) for k in keys) for s in pairs]
    hpairs = [tuple(_pin(s[k]) for k in keys) for s in pairs]
    it = [0]

    def one(src):
        def f():
            r = v.register(*src[it[0] % 3], **kw)
            it[0] += 1
            return r
        return f
    ms_d = _time(one(dpairs), 12)
    ms_h = _time(one(hpairs), 12)
    # parity + CPU time on pair 0
    s = pairs[0]
    r = v.register(*dpairs[0], **kw)
    t0 = time.perf_counter()
    m = cref.match_nn(s["scan_feat"], s["map_feat"])
    corr = match.filter_correspondences(m["idx01"], m["sim01"], min_cos=0.8)
    c = cref.ransac(s["scan_xyz"], s["map_xyz"], corr, None, 1.0, seed=1, n_hyp=1024)
    t_cpu = time.perf_counter() - t0
    rte, rre = synth.pose_errors(r.T, s["T_gt"])
    return {"workload": "configs[0]: 4096 x 4096 pts, 384-d, cos >= 0.8 gate, 1024 RANSAC hyps (tau = 1 m)",
            "value": 1e3 / ms_d, "e2e": 1e3 / ms_h, "unit": "pairs/s", "ms_per_pair": ms_d, "l2": "flushed between iterations",
            "h2d_bytes": 2 * 4096 * (384 + 3) * 4, "dtype": "f32 match / f64 solve",
            "cpu_baseline": {"value": 1.0 / t_cpu, "unit": "pairs/s", "cores": cref.num_threads(), "kind": "port"},
            "parity": {"corr_equal": bool(np.array_equal(r.corr, corr)), "mask_equal": bool(np.array_equal(r.inlier_mask, c["mask"])),
                       "best_equal": r.best_hyp == c["best"], "T_frob": float(np.linalg.norm(r.T - c["T"]))},
            "recall_at_1m_5deg": float(rte < 1.0 and rre < 5.0)}


def config3(v, ctx, dev, peaks):
    from oracle import cref, match
    from vfm_registration_b200 import synth
    n, m, d, h = 20_000, 200_000, 768, 65_536
    s = synth.make_pair(4, m, n, d, sigma_f=0.02)
    keys = ("scan_xyz", "map_xyz", "scan_feat", "map_feat")
    dpair = tuple(torch.from_numpy(s[k]).to(dev) for k in keys)
    kw = dict(min_cos=None, ratio=0.9, ransac_iters=h, inlier_thresh=1.0, seed=4)
    ctx.enable_timing(True)
    ms_d = _time(lambda: v.register(*dpair, **kw), 5, warmup=2, flush=False)   # 676 MB of inputs > L2
    k_ms, k_n = ctx.group_time_ms(0)
    r_ms, r_n = ctx.group_time_ms(1)
    ctx.enable_timing(False)
    hpair = tuple(_pin(s[k]) for k in keys)
    ms_h = _time(lambda: v.register(*hpair, **kw), 3, warmup=1, flush=False)
    r = v.register(*dpair, **kw)
    # parity on a bounded sample: the search is independent per query row, so the first 1500 rows against the full map
    rows = 1500
    t0 = time.perf_counter()
    c = cref.match_nn(s["scan_feat"][:rows], s["map_feat"])
    t_rows = time.perf_counter() - t0
    corr_head = match.filter_correspondences(c["idx01"], c["sim01"], c["sec01"], ratio=0.9)
    got = r.corr.cpu().numpy() if isinstance(r.corr, torch.Tensor) else r.corr
    flops = 2.0 * n * m * d
    k_avg = k_ms / max(k_n, 1)
    rte, rre = synth.pose_errors(r.T, s["T_gt"])
    return {"workload": "configs[3]: 200k map x 20k scan pts, 768-d, ratio test 0.9, 65536 RANSAC hyps (tau = 1 m)",
            "value": 1e3 / ms_d, "e2e": 1e3 / ms_h, "unit": "pairs/s", "ms_per_pair": ms_d, "hyps_per_sec": h * 1e3 / ms_d,
            "l2": "inputs (676 MB) larger than L2", "h2d_bytes": (n + m) * (d + 3) * 4, "dtype": "f32 match / f64 solve",
            "roofline": {"bound": "tensor", "kernel": "match_tc3_kernel<streamed A>, top-2 mode", "achieved": flops / (k_avg * 1e-3) / 1e12,
                         "peak": peaks["tf"], "unit": "TFLOP/s", "frac": flops / (k_avg * 1e-3) / 1e12 / peaks["tf"], "avg_launch_ms": k_avg,
                         "ransac_score_avg_ms": r_ms / max(r_n, 1),
                         "hyp_corr_per_sec": h * float(len(got)) / (r_ms / max(r_n, 1) * 1e-3)},
            "cpu_baseline": {"value": 1.0 / (t_rows * n / rows), "unit": "pairs/s (search only, scaled from the sample)",
                             "cores": cref.num_threads(), "kind": "port", "sample": f"{rows} of {n} scan rows against the full map"},
            "parity": {"against": f"oracle/c on the first {rows} scan rows (the full size is a -m gpu test)",
                       "corr_head_equal": bool(np.array_equal(got[got[:, 0] < rows], corr_head))},
            "n_corr": int(len(got)), "recall_at_1m_5deg": float(rte < 1.0 and rre < 5.0)}


def config2(v, ctx, dev, peaks):
    from scipy.spatial.transform import Rotation as R
    from vfm_registration_b200 import synth
    rng = np.random.default_rng(3)
    b, hh, ww, n_map, n_scan, model = 6, 224, 224, 50_000, 10_000, "vitl14"
    imgs = rng.integers(1, 255, (b, hh, ww, 3), dtype=np.uint8)
    imgs[0, 60:90, 40:80] = 0   # a black rectangle: exercises reject_black
    kmat = np.array([[200.0, 0, 112.0], [0, 200.0, 112.0], [0, 0, 1.0]])
    ts = []
    for i in range(b):
        t = np.eye(4)
        t[:3, :3] = (R.from_euler("z", 60.0 * i, degrees=True) * R.from_euler("yx", [90, -90], degrees=True)).as_matrix().T
        ts.append(t)
    ks, ts = np.stack([kmat] * b), np.stack(ts)
    map_xyz = np.c_[rng.uniform(-20, 20, (n_map, 2)), rng.uniform(-2, 4, n_map)].astype(np.float32)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        f = v.ViTFeaturizer(model, seed=4, random_init=True)   # no pretrained weights offline: architecture-level workload
    map_desc = v.extract_features(imgs, map_xyz, ks, ts, featurizer=f)
    resident = v.ResidentMap(torch.from_numpy(map_xyz).to(dev), map_desc)
    sel = rng.permutation(n_map)[:n_scan]
    t_gt = synth.random_pose(rng)
    t_inv = np.linalg.inv(t_gt)
    # the scan is the same physical points seen by the same rig (so its descriptors come from the same pixels), expressed
    # in the moved sensor frame for the solve
    scan_world = map_xyz[sel]
    scan_xyz = (scan_world.astype(np.float64) @ t_inv[:3, :3].T + t_inv[:3, 3] + rng.normal(0, 0.01, (n_scan, 3))).astype(np.float32)
    kw = dict(min_cos=0.8, mutual=True, ransac_iters=8192, inlier_thresh=0.5, seed=3)
    imgs_d, world_d, scan_d = torch.from_numpy(imgs).to(dev), torch.from_numpy(scan_world).to(dev), torch.from_numpy(scan_xyz).to(dev)
    imgs_h, world_h, scan_h = _pin(imgs), _pin(scan_world), _pin(scan_xyz)

    def run(im, pw, ps):
        desc = v.extract_features(im, pw, ks, ts, featurizer=f)
        if not isinstance(ps, torch.Tensor):
            ps = torch.from_numpy(ps).to(dev, non_blocking=True)
        return v.register_scans(resident, [(ps, desc)], **kw)[0]
    ctx.enable_timing(True)
    ms_d = _time(lambda: run(imgs_d, world_d, scan_d), 10)
    g_ms, g_n = ctx.group_time_ms(3)
    ctx.enable_timing(False)
    ms_vit = _time(lambda: f.forward(imgs_d), 10)
    ms_h = _time(lambda: run(imgs_h, world_h, scan_h), 10)
    r = run(imgs_d, world_d, scan_d)
    rte, rre = synth.pose_errors(r.T, t_gt)
    depth, width, tok = f.depth, f.width, 257
    flops_img = depth * (24.0 * tok * width * width + 4.0 * tok * tok * width) + 2.0 * 256 * 588 * width
    out = {"workload": "configs[2]: 6 x (224 x 224) images -> DINOv2 ViT-L/14 -> projection gather (10k pts) -> mutual match vs a "
                       "resident 50k-pt map -> 8192-hyp RANSAC; random-init weights (none available offline)",
           "value": 1e3 / ms_d, "e2e": 1e3 / ms_h, "unit": "pairs/s", "ms_per_pair": ms_d, "l2": "flushed between iterations",
           "vit_forward_ms": ms_vit, "images_per_sec": b * 1e3 / ms_vit, "h2d_bytes": b * hh * ww * 3 + 2 * n_scan * 12,
           "dtype": "bf16 GEMM operands, f32 accumulate / residual (ViT); f32 match / f64 solve",
           "roofline": {"bound": "tensor", "kernel": "vit_gemm_kernel (all GEMMs of one forward)", "achieved": b * flops_img / (ms_vit * 1e-3) / 1e12,
                        "peak": peaks["tf"], "unit": "TFLOP/s", "frac": b * flops_img / (ms_vit * 1e-3) / 1e12 / peaks["tf"],
                        "note": "whole forward (GEMMs + attention + norms) over the model's algorithmic flop",
                        "gemm_ms_per_forward": g_ms / max(g_n, 1) if g_n else None},
           "n_corr": int(len(r.corr)), "rte_m": rte, "rre_deg": rre}
    try:   # CPU side of this configuration: the fp32 torch restatement of the network (oracle/vit.py), one image
        from oracle import vit as ovit
        cfg = ovit.CONFIGS[model]
        sd = ovit.make_weights(cfg, seed=4)
        x = torch.stack([ovit.preprocess(imgs[0])])
        ovit.forward(sd, cfg, x)
        t0 = time.perf_counter()
        ovit.forward(sd, cfg, x)
        t1 = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 1.0 / (t1 * b), "unit": "pairs/s (ViT forward only: 6 images scaled from 1)", "cores": torch.get_num_threads(),
                               "kind": "port", "sample": "1 of 6 images through the torch fp32 restatement of ViT-L/14"}
    except Exception as e:  # pragma: no cover
        out["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
    resident.close()
    return out


def reference_shape(v, ctx, dev, peaks):
    from oracle import cref
    rng = np.random.default_rng(77)
    n, m, d = 300, 200_000, 384
    b = rng.standard_normal((m, d)).astype(np.float32)
    a = rng.standard_normal((n, d)).astype(np.float32)
    a[:120] = b[rng.integers(0, m, 120)] + 0.02 * rng.standard_normal((120, d)).astype(np.float32)
    resident = v.ResidentMap(torch.from_numpy(np.zeros((m, 3), np.float32)).to(dev), torch.from_numpy(b).to(dev))
    a_d = torch.from_numpy(a).to(dev)
    ctx.enable_timing(True)
    ms = _time(lambda: resident.match(a_d, min_cos=0.8, second=False), 20)   # the map (154 MB fp16) exceeds the L2; flushed anyway
    k_ms, k_n = ctx.group_time_ms(0)
    ctx.enable_timing(False)
    g = resident.match(a_d)
    c = cref.match_nn(a, b)
    k_avg = k_ms / max(k_n, 1)
    alg = 2.0 * d * m   # the fp16 map rows the search must read once
    resident.close()
    return {"workload": "reference shape (SURVEY D7): 300 queries x resident 200k-point map x 384-d, cosine gate 0.8",
            "call_ms": ms, "searches_per_sec": 1e3 / ms,
            "roofline": {"bound": "hbm", "kernel": "match_tc3_kernel on 300 x 200k x 384", "achieved": alg / (k_avg * 1e-3) / 1e9,
                         "peak": peaks["hbm"], "unit": "GB/s", "frac": alg / (k_avg * 1e-3) / 1e9 / peaks["hbm"], "avg_launch_ms": k_avg,
                         "algorithmic_bytes": alg},
            "parity": {"idx_equal": bool(np.array_equal(g.idx01.cpu().numpy(), c["idx01"])),
                       "sim_equal": bool(np.array_equal(g.sim01.cpu().numpy(), c["sim01"]))}}


LEGS = {"configs[0]": config0, "configs[3]": config3, "configs[2]": config2, "reference_shape": reference_shape}


def run_all(v, ctx, dev, peaks, names=None):
    out = {}
    for name in (names or LEGS):
        try:
            t0 = time.perf_counter()
            out[name] = LEGS[name](v, ctx, dev, peaks)
            out[name]["wall_s"] = time.perf_counter() - t0
        except Exception as e:
            out[name] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    import vfm_registration_b200 as v
    sys.path.insert(0, ROOT)
    from bench import load_peaks
    dev = torch.device("cuda", 0)
    print(json.dumps(run_all(v, v.get_context(0), dev, load_peaks(), sys.argv[1:] or None)))
