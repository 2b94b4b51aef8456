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


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "match"
    if what == "match":
        bench_match()
