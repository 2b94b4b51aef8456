"""How much does a small kernel cost the candidate search when both share the GPU?  Two contexts on one device, one stream
each: A = forward search (top-1 + gate) of a 10k x 50k x 384 pair in a loop, B = one kind of small work in a loop."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfm_registration_b200 as v
from vfm_registration_b200 import synth, api

dev = torch.device("cuda", 0)
s = synth.make_pair(2, 50_000, 10_000, 384)
ca, cb = api.Context(0), api.Context(0)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
import vfm_registration_b200.api as A
rm = None
def with_ctx(c, fn):
    A._contexts[0] = c
    return fn()
a = torch.from_numpy(s["scan_feat"]).to(dev)
with torch.cuda.stream(sa):
    rm = with_ctx(ca, lambda: v.ResidentMap(torch.from_numpy(s["map_xyz"]).to(dev), torch.from_numpy(s["map_feat"]).to(dev)))
m = with_ctx(ca, lambda: rm.match(a, min_cos=0.8, second=False))
corr = with_ctx(ca, lambda: v.filter_correspondences(m, min_cos=0.8))
sx, mx = torch.from_numpy(s["scan_xyz"]).to(dev), torch.from_numpy(s["map_xyz"]).to(dev)
big = torch.randn(50_000, 384, device=dev)
import ctypes as C
from vfm_registration_b200 import _lib
def search():
    with torch.cuda.stream(sa):
        with_ctx(ca, lambda: rm.match(a, min_cos=0.8, second=False))
k = corr.shape[0]
count = torch.full((1,), k, dtype=torch.int32, device=dev); T = torch.empty(16, dtype=torch.float64, device=dev)
mask = torch.zeros(k, dtype=torch.uint8, device=dev); stats = torch.zeros(4, dtype=torch.int64, device=dev)
def ransac():
    with torch.cuda.stream(sb):
        cb.bind_stream()
        _lib.check(cb.lib.vfmreg_ransac(cb.handle, A._ptr(sx), A._ptr(mx), 0, A._ptr(corr), A._ptr(count), k, None, 8192, 42, 1.0, 0,
                                       A._ptr(T), None, None, A._ptr(mask), A._ptr(stats)))
def normalize():
    with torch.cuda.stream(sb):
        with_ctx(cb, lambda: v.match_nn(big[:64], big, algo="tc"))   # dominated by renormalising the 50k x 384 'map'
def timeit(fa, fb, n=30):
    for f in (fa, fb):
        if f: f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        if fa: fa()
        if fb: fb()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6
print(f"search alone        {timeit(search, None):7.1f} us per iteration")
print(f"ransac alone        {timeit(None, ransac):7.1f}")
print(f"search + ransac     {timeit(search, ransac):7.1f}  (concurrent streams)")
print(f"normalize+small search alone {timeit(None, normalize):7.1f}")
print(f"search + that       {timeit(search, normalize):7.1f}")
