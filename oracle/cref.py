"""ctypes loader for the C oracle (``oracle/c/oracle_ref.c``) -- test infrastructure only.

Builds ``oracle/_build/liboracle.so`` with ``make -C oracle`` on first use when it is
missing (gcc is part of the image, both in the build container and on the GPU box)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, "c", "oracle_ref.c")):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        try:
            _lib = C.CDLL(_SO)
        except OSError:
            build(force=True)
            _lib = C.CDLL(_SO)
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_simd.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP threads of the C oracle (overrides OMP_NUM_THREADS, which torchrun sets to 1 for its workers)."""
    lib().orc_set_num_threads(C.c_int(int(n)))
    return num_threads()


def use_all_cores() -> int:
    """Run on every core the process is allowed on; returns the thread count."""
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        n = os.cpu_count() or 1
    return set_num_threads(n)


def renorm_l2(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    lib().orc_renorm_l2(_p(x, C.c_float), C.c_int64(x.shape[0]), C.c_int(x.shape[1]), _p(out, C.c_float))
    return out


def match_top2(a: np.ndarray, b: np.ndarray):
    """Canonical-order float32 inner-product top-2: returns (idx i32, best f32, second f32)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    n, d = a.shape
    m = b.shape[0]
    idx = np.empty(n, dtype=np.int32)
    best = np.empty(n, dtype=np.float32)
    second = np.empty(n, dtype=np.float32)
    lib().orc_match_top2(_p(a, C.c_float), C.c_int64(n), _p(b, C.c_float), C.c_int64(m), C.c_int(d),
                         _p(idx, C.c_int32), _p(best, C.c_float), _p(second, C.c_float))
    return idx, best, second


def match_nn(a, b, normalize=True, mutual=False):
    af = renorm_l2(a) if normalize else np.ascontiguousarray(a, dtype=np.float32)
    bf = renorm_l2(b) if normalize else np.ascontiguousarray(b, dtype=np.float32)
    out = {}
    out["idx01"], out["sim01"], out["sec01"] = match_top2(af, bf)
    if mutual:
        out["idx10"], out["sim10"], out["sec10"] = match_top2(bf, af)
    return out


def sample_indices(seed: int, n_hyp: int, n_corr: int) -> np.ndarray:
    out = np.empty((n_hyp, 3), dtype=np.int32)
    lib().orc_sample_indices(C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), C.c_int64(n_hyp), C.c_int32(n_corr),
                             _p(out, C.c_int32))
    return out


def kabsch3(p: np.ndarray, q: np.ndarray):
    p = np.ascontiguousarray(p, dtype=np.float64).reshape(9)
    q = np.ascontiguousarray(q, dtype=np.float64).reshape(9)
    rt = np.empty(12, dtype=np.float64)
    lib().orc_kabsch3.restype = C.c_int
    ok = lib().orc_kabsch3(_p(p, C.c_double), _p(q, C.c_double), _p(rt, C.c_double))
    return rt[:9].reshape(3, 3).copy(), rt[9:].copy(), bool(ok)


def ransac(src_xyz, tgt_xyz, corr, sample_idx, thresh, refit=False, seed=0, n_hyp=None):
    src = np.ascontiguousarray(src_xyz, dtype=np.float64)
    tgt = np.ascontiguousarray(tgt_xyz, dtype=np.float64)
    corr = np.ascontiguousarray(np.asarray(corr).reshape(-1, 2), dtype=np.int32)
    k = corr.shape[0]
    pq = np.empty((max(k, 1), 6), dtype=np.float64)
    if k:
        lib().orc_gather_pq(_p(src, C.c_double), _p(tgt, C.c_double), _p(corr, C.c_int32), C.c_int32(k),
                            _p(pq, C.c_double))
    if sample_idx is not None:
        sample_idx = np.ascontiguousarray(sample_idx, dtype=np.int32)
        h = sample_idx.shape[0]
    else:
        h = int(n_hyp)
    t = np.empty(16, dtype=np.float64)
    counts = np.empty(max(h, 1), dtype=np.int32)
    sumq = np.empty(max(h, 1), dtype=np.int64)
    mask = np.zeros(max(k, 1), dtype=np.uint8)
    stats = np.zeros(4, dtype=np.int64)
    lib().orc_ransac(_p(pq, C.c_double), C.c_int32(k), _p(sample_idx, C.c_int32), C.c_int64(h),
                     C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), C.c_double(thresh), C.c_int(int(refit)),
                     _p(t, C.c_double), _p(counts, C.c_int32), _p(sumq, C.c_int64), _p(mask, C.c_uint8),
                     _p(stats, C.c_int64))
    n_in = int(stats[1])
    tau2 = float(thresh) ** 2
    rmse = float(np.sqrt(stats[2] / 2.0 ** 40 * tau2 / n_in)) if n_in else 0.0
    return dict(T=t.reshape(4, 4), best=int(stats[0]), counts=counts[:h], sumq=sumq[:h],
                mask=mask[:k].astype(bool), fitness=(n_in / k if k else 0.0), rmse=rmse, n_inliers=n_in)
