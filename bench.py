#!/usr/bin/env python
"""bench.py -- scene-pairs/s of the descriptor-match-and-solve hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scenes-per-step S] [--scans-per-map C]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] -- NCLT-shape pairs, 50k map x 10k scan points, 384-d descriptors,
mutual nearest neighbour + cosine gate 0.8, 8192 RANSAC hypotheses (tau = 1 m) -- laid out as the reference lays its scenes
out: one local map per scene and C = 5 scans registered against it (registration_node.py:554-590, data/nclt/scene_*.json).
A "step" is one pass of the hot path over S scenes = S x C pairs of synthetic input (S x 77 MB of maps + S x C x 15 MB of scans
per step > the 126 MB L2, so consecutive scenes evict each other); weak scaling: every rank owns its own S scenes and the
per-pair 4x4 transforms are all-gathered once per step.

  value  = pairs/s with the inputs already resident in HBM (register_batch on CUDA tensors -> vfmreg_register_batch:
           consecutive pairs on `--lanes` streams, the candidate-search kernels on two high-priority streams, every map
           prepared once for its C scans)
  e2e    = pairs/s through the public API (register_batch) with HOST (pinned) buffers, H2D + D2H inside the timed region;
           `e2e.pageable` = the same with plain (pageable) NumPy arrays, as a reference user would pass them
  distinct = the round-1 layout for comparison: every pair brings its own map (nothing shared)
  roofline = the dominant kernel (descriptor N x M candidate search, scan -> map), timed with CUDA events on the stream it
             is launched on, inside the timed region (`frac`) and once more with nothing beside it (`alone`)
  parity = the last scene's results against the oracle (oracle/c) on the same inputs: correspondence lists, inlier masks and
           winning hypotheses equal, max Frobenius distance of the transforms
  cpu_baseline = oracle/c (the reference-style CPU restatement; the reference itself cannot be built offline) on the
                 box's host cores, one scene.  `--impl reference` times that same CPU path as the reference arm.
  extras = short measurements of BASELINE configs[0], configs[2] and configs[3] (N = 1 only; --no-extras skips them)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_MAP, N_SCAN, DIM, N_HYP = 50_000, 10_000, 384, int(os.environ.get("VFM_BENCH_HYPS", "8192"))   # env: tuning experiments only
MIN_COS, TAU = 0.8, 1.0
METRIC, UNIT = "scene_pairs_per_sec", "pairs/s"
WORKLOAD = ("configs[1]: NCLT-shape pairs, 50k map x 10k scan pts, 384-d feats, mutual-NN + cos>=0.8 gate, "
            "8192 RANSAC hyps (tau=1m); scenes of one map + 5 scans as in the reference (registration_node.py:554-590)")
RECALL_THRESHOLDS = ((1.0, 5.0), (0.3, 15.0), (0.6, 1.5), (2.0, 5.0))   # BASELINE's + the reference's (registration_node.py:973-977)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "10",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        busy = sorted(sm)[len(sm) // 2:] or [0.0]  # upper half of the samples = under load
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(index: int):
    """One process per GPU: run on the cores NVML reports as local to that GPU, so that the pinned host buffers of the e2e leg
    are first-touched on the GPU's NUMA node (a rank left on the far socket halves its H2D rate).  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"{len(allowed)} cores local to GPU {index}"
    except Exception as e:  # pragma: no cover
        return f"not bound ({type(e).__name__})"
    return "not bound"


# ---- the reference-style CPU path (oracle/c): the checker and the timed CPU baseline ------------------------------------
def cpu_scene(scene, n_hyp=None):
    """One scene on the host cores, structured as the reference structures it: the map is renormalised once, every scan is
    searched in both directions (find_correspondences' two queries, registration_node.py:487-527), gated, and solved.
    Returns (seconds, [result dict per scan])."""
    from oracle import cref, match
    n_hyp = n_hyp or N_HYP
    t0 = time.perf_counter()
    mf = cref.renorm_l2(scene["map_feat"])
    out = []
    for sc in scene["scans"]:
        sf = cref.renorm_l2(sc["scan_feat"])
        i01, s01, c01 = cref.match_top2(sf, mf)
        i10, _, _ = cref.match_top2(mf, sf)
        corr = match.filter_correspondences(i01, s01, c01, i10, min_cos=MIN_COS, mutual=True)
        r = cref.ransac(sc["scan_xyz"], scene["map_xyz"], corr, None, TAU, seed=42, n_hyp=n_hyp)
        r["corr"] = corr
        out.append(r)
    return time.perf_counter() - t0, out


def recall_table(errs):
    return {f"{t:g}m_{r:g}deg": float(np.mean([(e[0] < t) and (e[1] < r) for e in errs])) for t, r in RECALL_THRESHOLDS}


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path.  Open3D / faiss / kiss_icp cannot be installed
    offline, so this is the oracle port (kind 'port') on ALL host cores the process may use (torchrun's
    OMP_NUM_THREADS=1 is overridden), one full scene (1 map + C scans, nothing sub-sampled) per step."""
    if rank != 0:
        return
    from oracle import cref
    from vfm_registration_b200 import synth
    try:
        os.sched_setaffinity(0, range(os.cpu_count()))   # a launcher may have pinned this rank to a few cores
    except (AttributeError, OSError):
        pass
    cores = cref.use_all_cores()
    scene = synth.make_scene(1000, N_MAP, args.scans_per_map, N_SCAN, DIM)
    times, errs = [], []
    for i in range(args.warmup + args.steps):
        t, res = cpu_scene(scene)
        if i >= args.warmup:
            times.append(t)
            errs = [synth.pose_errors(r["T"], sc["T_gt"]) for r, sc in zip(res, scene["scans"])]
    per_step = float(np.mean(times))
    val = args.scans_per_map / per_step
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 match / f64 solve",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "pairs_per_step": args.scans_per_map},
            "recall": recall_table(errs),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"1 full scene per step: 1 map (50k x 384) renormalised once + {args.scans_per_map} scans, each "
                                       f"matched against the full map in both directions and solved with {N_HYP} hypotheses; "
                                       f"oracle/c restatement (AVX2 + OpenMP, {cores} threads)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_other_workload(args, v, ctx, dev, rank, local_rank):
    """`--workload configs[0|2|3]` / `reference_shape`: one of the measurement legs of tools/bench_extras.py as a bench line of its
    own (rank 0, one GPU): value = device-resident, e2e = host buffers, roofline of the leg's dominant kernel, parity and the CPU
    arm of that configuration where the leg has them."""
    if rank != 0:
        return
    from tools import bench_extras
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.kernel_launches
    leg = bench_extras.LEGS[args.workload](v, ctx, dev, load_peaks())
    launches = ctx.kernel_launches - l0
    clocks = sampler.stop()
    unit = "searches/s" if args.workload == "reference_shape" else UNIT
    value = leg.get("value", leg.get("searches_per_sec"))
    line = {"metric": METRIC if unit == UNIT else "map_searches_per_sec", "value": value, "unit": unit, "n_gpus": 1, "steps": 10, "warmup": 3,
            "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": leg.get("dtype", "f32 match"), "data": "synthetic", "config": {"workload": leg["workload"], "l2": leg.get("l2")},
            "roofline": leg.get("roofline"), "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": leg.get("e2e"), "unit": unit, "h2d_bytes_per_step": leg.get("h2d_bytes"), "d2h_bytes_per_step": None},
            "parity": leg.get("parity"), "cpu_baseline": leg.get("cpu_baseline")}
    for k in ("hyps_per_sec", "recall_at_1m_5deg", "vit_forward_ms", "images_per_sec", "n_corr", "rte_m", "rre_deg"):
        if k in leg:
            line[k] = leg[k]
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes-per-step", type=int, default=4)
    ap.add_argument("--scans-per-map", type=int, default=5, help="5 = NCLT scenes, 3 = RobotCar scenes")
    ap.add_argument("--distinct-pairs", type=int, default=8, help="pairs of the `distinct` comparison leg (0 = skip)")
    ap.add_argument("--lanes", type=int, default=0, help="compute lanes of register_batch (0 = library default)")
    ap.add_argument("--algo", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--workload", default="configs[1]", choices=["configs[0]", "configs[1]", "configs[2]", "configs[3]", "reference_shape"],
                    help="configs[1] = the headline (BASELINE.json's metric is quoted on it); the others print the same JSON line "
                         "for one of the other BASELINE configurations (N = 1 only; they are parity-test cases first)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1 and os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
        # stdout carries exactly one JSON line: NCCL's log (banner, communicator lines) goes to stderr, at the level asked for
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import vfm_registration_b200 as v
    from vfm_registration_b200 import synth
    ctx = v.get_context(local_rank)
    if args.workload != "configs[1]":
        run_other_workload(args, v, ctx, dev, rank, local_rank)
        if world > 1:
            dist.destroy_process_group()
        return
    lanes = args.lanes or 5   # the library default
    ctx.set_lanes(lanes)
    S, C = args.scenes_per_step, args.scans_per_map
    P = S * C
    peaks = load_peaks()

    # ---- synthetic inputs: S scenes per rank (one map + C scans each), resident on the device AND in pinned host memory
    scenes = [synth.make_scene(1000 + rank * S + s, N_MAP, C, N_SCAN, DIM) for s in range(S)]
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()   # noqa: E731
    dev_pairs, pin_pairs, page_pairs, truth = [], [], [], []
    for sc in scenes:
        mx_d, mf_d = torch.from_numpy(sc["map_xyz"]).to(dev), torch.from_numpy(sc["map_feat"]).to(dev)
        mx_p, mf_p = pin(sc["map_xyz"]), pin(sc["map_feat"])
        for s in sc["scans"]:
            dev_pairs.append((torch.from_numpy(s["scan_xyz"]).to(dev), mx_d, torch.from_numpy(s["scan_feat"]).to(dev), mf_d))
            pin_pairs.append((pin(s["scan_xyz"]), mx_p, pin(s["scan_feat"]), mf_p))
            page_pairs.append((s["scan_xyz"], sc["map_xyz"], s["scan_feat"], sc["map_feat"]))
            truth.append(s["T_gt"])
    kw = dict(min_cos=MIN_COS, mutual=True, ransac_iters=N_HYP, inlier_thresh=TAU, seed=42, algo=args.algo)

    def make_step(n_pairs):
        t_all = torch.zeros((world * n_pairs, 4, 4), dtype=torch.float64, device=dev)
        t_loc = torch.zeros((n_pairs, 4, 4), dtype=torch.float64, device=dev)

        def step(inputs):
            # host buffers: the H2D of pair i+1 (and of the next map) overlaps the solve of pair i inside the library;
            # CUDA tensors: all pairs enqueued back to back (one synchronisation per step)
            res = v.register_batch(inputs, **kw)
            t_loc.copy_(torch.from_numpy(np.stack([r.T for r in res])))
            if world > 1:
                dist.all_gather_into_tensor(t_all, t_loc)  # the path's only collective: (P, 4, 4) transforms per rank
            return res
        return step

    def timed(inputs, steps, warmup):
        step = make_step(len(inputs))
        for _ in range(warmup):
            step(inputs)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.kernel_launches
        e0.record()
        for _ in range(steps):
            res = step(inputs)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), ctx.kernel_launches - l0, res

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.enable_timing(True)
    ms_dev, launches, res = timed(dev_pairs, args.steps, args.warmup)
    match_ms, match_launches = ctx.group_time_ms(0)
    ransac_ms, ransac_launches = ctx.group_time_ms(1)
    pruned_ms, pruned_launches = ctx.group_time_ms(4)
    ctx.enable_timing(False)
    # the same kernel timed with nothing beside it (one lane), for the record: inside the timed region above it shares
    # the SMs with the small kernels of the neighbouring pairs
    ctx.set_lanes(1)
    ctx.enable_timing(True)
    timed(dev_pairs, 2, 1)
    alone_ms, alone_launches = ctx.group_time_ms(0)
    ctx.enable_timing(False)
    ctx.set_lanes(lanes)
    e2e_steps = max(2, args.steps // 2)
    ms_e2e, _, res_e2e = timed(pin_pairs, e2e_steps, 2)
    clocks = sampler.stop() if rank == 0 else None   # sampled over the device-resident and the host-buffer timed regions
    page_steps = 2
    ms_page, _, res_page = timed(page_pairs, page_steps, 1)

    # the three routes agree bit for bit
    for a, b, c in zip(res, res_e2e, res_page):
        assert np.array_equal(a.T, b.T) and np.array_equal(a.T, c.T), "device-pointer, pinned-host and pageable-host paths disagree"
        assert np.array_equal(a.corr.cpu().numpy(), b.corr) and np.array_equal(b.corr, c.corr)
    errs = [synth.pose_errors(r.T, t) for r, t in zip(res, truth)]

    # ---- the round-1 layout for comparison: every pair brings its own map
    distinct = None
    if args.distinct_pairs > 0:
        D = args.distinct_pairs
        dd, dp = [], []
        for p in range(D):
            s = synth.make_pair(5000 + rank * D + p, N_MAP, N_SCAN, DIM)
            keys = ("scan_xyz", "map_xyz", "scan_feat", "map_feat")
            dd.append(tuple(torch.from_numpy(s[k]).to(dev) for k in keys))
            dp.append(tuple(pin(s[k]) for k in keys))
        ms_dd, _, _ = timed(dd, max(2, args.steps // 2), 2)
        ms_dp, _, _ = timed(dp, 2, 1)
        distinct = {"pairs_per_step_per_gpu": D, "value": world * D * max(2, args.steps // 2) / (ms_dd / 1e3),
                    "e2e": world * D * 2 / (ms_dp / 1e3), "h2d_bytes_per_step": int(sum(x.nbytes for pp in dp for x in pp)), "unit": UNIT}
        del dd, dp

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_pairs = world * P * args.steps
    value = total_pairs / (ms_dev / 1e3)
    e2e_val = world * P * e2e_steps / (ms_e2e / 1e3)
    # roofline of the dominant kernel: the full scan->map search, 2*N*M*D flop per launch.  The map->scan direction of the
    # mutual check only needs idx10[j] at the map rows j that gated queries point at, so register() searches those
    # (<= N) rows only (`pruned_*` below; identical correspondences, tests/test_gpu_parity.py::test_pruned_mutual_*).
    flop_per_launch = 2.0 * N_SCAN * N_MAP * DIM
    alg_bytes_per_launch = 4.0 * DIM * (N_SCAN + N_MAP)
    # warm-up launches are included in the event total, so divide by the launches actually recorded
    avg_ms = match_ms / max(match_launches, 1)
    achieved = flop_per_launch / (avg_ms * 1e-3) / 1e12
    alone_avg = alone_ms / max(alone_launches, 1)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_match_tc.json")
    if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel
        tj = json.load(open(tpath))
        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    n_timed = (args.steps + args.warmup) * P
    roofline = {"bound": "tensor", "kernel": "match_tc3_kernel: descriptor N x M candidate search, scan -> map (full), timed while sharing the SMs with the other lanes",
                "achieved": achieved, "peak": peaks["tf"], "unit": "TFLOP/s", "frac": achieved / peaks["tf"],
                "peak_source": f"{peaks['source']} bf16 burst", "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write)",
                "algorithmic_bytes": alg_bytes_per_launch, "avg_launch_ms": avg_ms,
                "alone": {"avg_launch_ms": alone_avg, "achieved": flop_per_launch / (alone_avg * 1e-3) / 1e12,
                          "frac": flop_per_launch / (alone_avg * 1e-3) / 1e12 / peaks["tf"],
                          "note": "same kernel, one lane: nothing else resident on the SMs"},
                "launches_timed": match_launches, "algorithmic_gbs": alg_bytes_per_launch / (avg_ms * 1e-3) / 1e9,
                "share_of_step": avg_ms * (match_launches / n_timed) * P / (ms_dev / args.steps),
                "full_search_launches_per_pair": match_launches / n_timed,
                "pruned_reverse_search_avg_ms": pruned_ms / max(pruned_launches, 1),
                "pruned_reverse_search_launches_per_pair": pruned_launches / n_timed,
                "ransac_score_avg_ms": ransac_ms / max(ransac_launches, 1)}
    nbytes = lambda t: int(sum(x.nbytes for x in t))   # noqa: E731
    # bytes the library copies per step: every map once (xyz + descriptors), every scan once
    h2d = sum(nbytes((pp[1], pp[3])) for pp in pin_pairs[::C]) + sum(nbytes((pp[0], pp[2])) for pp in pin_pairs)
    d2h = P * (N_SCAN * 2 * 4 + N_SCAN + 256)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 match / f64 solve", "data": "synthetic",
            "config": {"workload": WORKLOAD, "scenes_per_step_per_gpu": S, "scans_per_map": C, "pairs_per_step_per_gpu": P,
                       "parallelism": f"scenes sharded over {world} rank(s)",
                       "l2": f"{S} maps x 77 MB + {P} scans x 15 MB cycled per step (> 126 MB L2)", "algo": args.algo,
                       "lanes": lanes, "host_affinity": numa},
            "hyps_per_sec": value * N_HYP, "recall_at_1m_5deg": recall_table(errs)["1m_5deg"], "recall": recall_table(errs),
            "roofline": roofline, "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / e2e_steps, "host_memory": "pinned",
                    "pageable": {"value": world * P * page_steps / (ms_page / 1e3), "unit": UNIT,
                                 "note": "plain NumPy arrays (what a reference user passes): staged through pinned memory by the library's copy threads (csrc/hostcopy.cu)"}},
            "distinct": distinct,
            "gpu_launches": int(launches)}
    # ---- parity against the oracle + the CPU baseline, on the LAST scene of this rank (one full scene on the host cores)
    if not args.no_cpu_baseline:
        from oracle import cref
        try:
            os.sched_setaffinity(0, range(os.cpu_count()))
        except (AttributeError, OSError):
            pass
        cores = cref.use_all_cores()   # torchrun exports OMP_NUM_THREADS=1; the baseline uses every core it may
        t_cpu, cres = cpu_scene(scenes[-1])
        n_par = C
        if world == 1 and S >= 2:   # a second scene: parity and recall-vs-CPU on 2 x C >= 8 pairs (N = 1 only: 3 s of host time)
            _, cres0 = cpu_scene(scenes[-2])
            cres, n_par = cres0 + cres, 2 * C
        gres = res[-n_par:]
        t_frob = max(float(np.linalg.norm(g.T - c["T"])) for g, c in zip(gres, cres))
        line["parity"] = {"against": "oracle/c on the last scene(s) of the step (same inputs, same seed)", "pairs": n_par,
                          "corr_equal": all(np.array_equal(g.corr.cpu().numpy(), c["corr"]) for g, c in zip(gres, cres)),
                          "mask_equal": all(np.array_equal(g.inlier_mask.cpu().numpy(), c["mask"]) for g, c in zip(gres, cres)),
                          "best_equal": all(g.best_hyp == c["best"] for g, c in zip(gres, cres)),
                          "T_frob": t_frob}
        cpu_errs = [synth.pose_errors(c["T"], t) for c, t in zip(cres, truth[-n_par:])]
        line["parity"]["recall_equal"] = recall_table(cpu_errs) == recall_table(errs[-n_par:])
        if world == 1:
            line["cpu_baseline"] = {"value": C / t_cpu, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"1 full scene: map renormalised once + {C} scans x (10k x 50k x 384 search in both "
                                              f"directions + gate + {N_HYP}-hyp RANSAC) in {t_cpu:.2f} s; oracle/c restatement "
                                              f"(AVX2 + OpenMP)", "recall": recall_table(cpu_errs)}
    if world == 1 and not args.no_extras:
        try:
            from tools import bench_extras
            line["extras"] = bench_extras.run_all(v, ctx, dev, peaks)
        except Exception as e:   # the extras never take the headline down with them
            line["extras"] = {"error": f"{type(e).__name__}: {e}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
