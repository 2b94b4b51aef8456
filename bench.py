#!/usr/bin/env python
"""bench.py -- scene-pairs/s of the descriptor-match-and-solve hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs-per-step P]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] -- NCLT-shape pairs, 50k map x 10k scan points, 384-d descriptors,
mutual nearest neighbour + cosine gate 0.8, 8192 RANSAC hypotheses (tau = 1 m).  A "step" is one pass of the hot path
over a batch of P distinct synthetic pairs (P x 92 MB of inputs > the 126 MB L2, so consecutive pairs evict each
other); weak scaling: every rank owns its own P pairs and the per-pair 4x4 transforms are all-gathered once per step.

  value  = pairs/s with the inputs already resident in HBM (register_batch on CUDA tensors -> vfmreg_register_batch:
           consecutive pairs on `--lanes` streams, the candidate-search kernels on two high-priority streams)
  e2e    = pairs/s through the public API (register_batch) with HOST (pinned) buffers, H2D + D2H inside the timed region
  roofline = the dominant kernel (descriptor N x M candidate search, scan -> map), timed with CUDA events on the stream it
             is launched on, inside the timed region (`frac`) and once more with nothing beside it (`alone`)
  cpu_baseline = oracle/c (the reference-style CPU restatement; the reference itself cannot be built offline) on the
                 box's host cores, bounded sample.  `--impl reference` times that same CPU path as the reference arm.
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
WORKLOAD = ("configs[1]: NCLT-shape pair, 50k map x 10k scan pts, 384-d feats, mutual-NN + cos>=0.8 gate, "
            "8192 RANSAC hyps (tau=1m)")


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


def cpu_path(pair, n_scan_rows, threads_note=True):
    """The reference-style CPU path (oracle/c) on one pair; match restricted to the first n_scan_rows scan points for
    the scan->map direction and timed separately so it can be scaled."""
    from oracle import cref, match
    t0 = time.perf_counter()
    sf = cref.renorm_l2(pair["scan_feat"][:n_scan_rows])
    mf = cref.renorm_l2(pair["map_feat"])
    i01, s01, c01 = cref.match_top2(sf, mf)
    i10, s10, c10 = cref.match_top2(mf, sf)
    t_match = time.perf_counter() - t0
    corr = match.filter_correspondences(i01, s01, c01, i10, min_cos=MIN_COS, mutual=True)
    t1 = time.perf_counter()
    r = cref.ransac(pair["scan_xyz"], pair["map_xyz"], corr, None, TAU, seed=42, n_hyp=N_HYP)
    t_ransac = time.perf_counter() - t1
    return t_match, t_ransac, r, corr


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path.  Open3D / faiss / kiss_icp cannot be installed
    offline, so this is the oracle port (kind 'port'), all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import cref
    from vfm_registration_b200 import synth
    cores = cref.num_threads()
    pair = synth.make_pair(2, N_MAP, N_SCAN, DIM)
    # probe to size the per-step sample at <= ~6 s
    tm, tr, _, _ = cpu_path(pair, 250)
    per_row = tm / 250.0  # both directions scale with the scan rows
    rows = int(min(N_SCAN, max(250, 6.0 / max(per_row, 1e-9))))
    times = []
    for i in range(args.warmup + args.steps):
        tm, tr, r, corr = cpu_path(pair, rows)
        if i >= args.warmup:
            times.append(tm * (N_SCAN / rows) + tr)
    per_pair = float(np.mean(times))
    val = 1.0 / per_pair
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_pair * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "pairs_per_step": 1},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"1 pair/step, scan rows {rows}/{N_SCAN} matched against the full 50k map in both "
                                       f"directions (match time scaled x{N_SCAN / rows:.2f}), full 8192-hyp RANSAC"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-step", type=int, default=16)
    ap.add_argument("--lanes", type=int, default=0, help="compute lanes of register_batch (0 = library default)")
    ap.add_argument("--algo", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "WARN"   # NCCL prints its banner on stdout; stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)

    import vfm_registration_b200 as v
    from vfm_registration_b200 import synth
    ctx = v.get_context(local_rank)
    lanes = args.lanes or 5   # the library default
    ctx.set_lanes(lanes)
    P = args.pairs_per_step
    peaks = load_peaks()

    # ---- synthetic inputs: P distinct pairs per rank, resident on the device AND in pinned host memory
    pairs, dev_pairs, pin_pairs = [], [], []
    for p in range(P):
        s = synth.make_pair(1000 + rank * P + p, N_MAP, N_SCAN, DIM)
        pairs.append(s)
        keys = ("scan_xyz", "map_xyz", "scan_feat", "map_feat")
        dev_pairs.append(tuple(torch.from_numpy(s[k]).to(dev) for k in keys))
        pin_pairs.append(tuple(torch.from_numpy(s[k]).pin_memory().numpy() for k in keys))
    kw = dict(min_cos=MIN_COS, mutual=True, ransac_iters=N_HYP, inlier_thresh=TAU, seed=42, algo=args.algo)
    t_all = torch.zeros((world * P, 4, 4), dtype=torch.float64, device=dev)
    t_loc = torch.zeros((P, 4, 4), dtype=torch.float64, device=dev)

    def step(inputs, host=False):
        # host=True: pinned host buffers, H2D of pair i+1 overlaps the solve of pair i inside the library;
        # host=False: device-resident inputs, all pairs enqueued back to back (one synchronisation per step)
        res = v.register_batch(inputs, **kw)
        for p in range(P):
            t_loc[p].copy_(torch.from_numpy(res[p].T), non_blocking=False)
        if world > 1:
            dist.all_gather_into_tensor(t_all, t_loc)  # the path's only collective: (P, 4, 4) transforms per rank
        return res

    def timed(inputs, steps, warmup, host=False):
        for _ in range(warmup):
            step(inputs, host)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.kernel_launches
        e0.record()
        for _ in range(steps):
            res = step(inputs, host)
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
    ms_e2e, _, res_e2e = timed(pin_pairs, max(2, args.steps // 2), 2, host=True)
    e2e_steps = max(2, args.steps // 2)
    clocks = sampler.stop() if rank == 0 else None   # sampled over the device-resident and the host-buffer timed regions

    # parity guard inside the bench: device path and host path give the same transforms
    for a, b in zip(res, res_e2e):
        assert np.array_equal(a.T, b.T), "device-pointer and host-buffer paths disagree"
    errs = [synth.pose_errors(r.T, s["T_gt"]) for r, s in zip(res, pairs)]
    recall = float(np.mean([(e[0] < 1.0) and (e[1] < 5.0) for e in errs]))

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
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_match_tc.json")
    if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel
        tj = json.load(open(tpath))
        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    roofline = {"bound": "tensor", "kernel": "match_tc3_kernel: descriptor N x M candidate search, scan -> map (full), timed while sharing the SMs with the other lanes",
                "achieved": achieved, "peak": peaks["tf"], "unit": "TFLOP/s", "frac": achieved / peaks["tf"],
                "peak_source": f"{peaks['source']} bf16 burst", "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write)",
                "algorithmic_bytes": alg_bytes_per_launch, "avg_launch_ms": avg_ms,
                "alone": {"avg_launch_ms": alone_ms / max(alone_launches, 1),
                          "achieved": flop_per_launch / (alone_ms / max(alone_launches, 1) * 1e-3) / 1e12,
                          "frac": flop_per_launch / (alone_ms / max(alone_launches, 1) * 1e-3) / 1e12 / peaks["tf"],
                          "note": "same kernel, one lane: nothing else resident on the SMs"},
                "launches_timed": match_launches, "algorithmic_gbs": alg_bytes_per_launch / (avg_ms * 1e-3) / 1e9,
                "share_of_step": avg_ms * (match_launches / ((args.steps + args.warmup) * P)) * P / (ms_dev / args.steps),
                "full_search_launches_per_pair": match_launches / ((args.steps + args.warmup) * P),
                "pruned_reverse_search_avg_ms": pruned_ms / max(pruned_launches, 1),
                "pruned_reverse_search_launches_per_pair": pruned_launches / ((args.steps + args.warmup) * P),
                "ransac_score_avg_ms": ransac_ms / max(ransac_launches, 1)}
    nbytes = lambda t: int(sum(x.nbytes for x in t))
    h2d = sum(nbytes(pp) for pp in pin_pairs)
    d2h = P * (N_SCAN * 2 * 4 + N_SCAN + 168)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 match / f64 solve", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": P, "parallelism": f"pairs sharded over {world} rank(s)",
                       "l2": f"{P} distinct pairs x 92 MB cycled per step (> 126 MB L2)", "algo": args.algo,
                       "lanes": lanes, "host_affinity": numa},
            "hyps_per_sec": value * N_HYP, "recall_at_1m_5deg": recall,
            "roofline": roofline, "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / e2e_steps},
            "gpu_launches": int(launches)}
    if world == 1 and not args.no_cpu_baseline:
        from oracle import cref
        # bounded sample: ONE full pair on the host cores (~0.5-1 s with 16+ threads), best of 2
        best = None
        for _ in range(2):
            tm, tr, rc, _ = cpu_path(pairs[-1], N_SCAN)
            best = (tm, tr) if best is None or tm + tr < sum(best) else best
        tm, tr = best
        per_pair = tm + tr
        rte, rre = synth.pose_errors(rc["T"], pairs[-1]["T_gt"])
        line["cpu_baseline"] = {"value": 1.0 / per_pair, "unit": UNIT, "cores": cref.num_threads(), "kind": "port",
                                "sample": f"1 full pair (10k x 50k x 384, both directions: {tm:.2f}s; 8192-hyp RANSAC: {tr:.3f}s), "
                                          f"best of 2; oracle/c restatement (AVX2 + OpenMP)",
                                "recall_ok": bool(rte < 1.0 and rre < 5.0)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
