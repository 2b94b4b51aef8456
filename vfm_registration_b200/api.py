"""Python host of libvfmreg_b200.so: the call surface BASELINE.json names --
``extract_features()`` / ``register(source_pcd, target_pcd, src_feats, tgt_feats)`` -- plus the low-level ops.

PyTorch is used for device memory and streams only; every computation happens in the hand-written sm_100a
kernels behind the C ABI (include/vfmreg_b200.h).  NumPy inputs go through the host-buffer entry point
(``vfmreg_register_host``: host->device copies inside the call); CUDA tensors are borrowed zero-copy.

Reference call sites these functions replace (paths relative to the reference checkout):
  register             RegistrationNode.ransac_registration(method='vfm', run_icp=False)
                       src/vfm-reg/src/registration_node.py:273-328 (+ compute_vfm_correspondences :396-425)
  match_nn             VoxelHashMap::GetVFMCorrespondences' faiss block, VoxelHashMap.cpp:469-511
  ransac_kabsch        o3d registration_ransac_based_on_correspondence, registration_node.py:319-327
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import VfmRegError

__all__ = ["Context", "get_context", "match_nn", "filter_correspondences", "ransac_kabsch", "register", "RegResult",
           "MatchResult", "RansacResult", "VfmRegError", "CameraSpec", "project_gather", "register_batch", "ResidentMap",
           "register_scans", "l2_distances", "select_smallest", "KdTree", "ransac_nn_all", "RansacNNResult"]

_ALGO = {"auto": _lib.ALGO_AUTO, "simt": _lib.ALGO_SIMT, "tc": _lib.ALGO_TC}


class Context:
    """One per (process, device).  Owns the library context (scratch arena, stream binding)."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self.lib.vfmreg_create(int(device), C.byref(h)), "vfmreg_create")
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.vfmreg_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def bind_stream(self):
        s = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.vfmreg_set_stream(self.handle, C.c_void_p(s)), "vfmreg_set_stream")

    def sync(self):
        _lib.check(self.lib.vfmreg_sync(self.handle), "vfmreg_sync")

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.vfmreg_kernel_launches(self.handle))

    def enable_timing(self, on: bool = True):
        _lib.check(self.lib.vfmreg_enable_timing(self.handle, int(on)))

    def set_lanes(self, lanes: int):
        """Number of CUDA streams ``register_batch`` spreads consecutive (device-resident) pairs over (1..8, default 5)."""
        _lib.check(self.lib.vfmreg_set_lanes(self.handle, int(lanes)), "vfmreg_set_lanes")

    def group_time_ms(self, group: int):
        ms, n = C.c_float(), C.c_int()
        _lib.check(self.lib.vfmreg_group_time_ms(self.handle, group, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)


_contexts: dict = {}


def get_context(device: Optional[int] = None) -> Context:
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = _contexts[device] = Context(device)
    return ctx


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _dev_f32(x, device, name, cols=None) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{name}: expected ndarray or Tensor, got {type(x)}")
    if x.dim() != 2 or (cols is not None and x.shape[1] != cols):
        raise ValueError(f"Invalid shape for {name}: {tuple(x.shape)}")  # reference: mapping.py:72-73
    return x.to(device=device, dtype=torch.float32).contiguous()


@dataclass
class MatchResult:
    idx01: torch.Tensor
    sim01: torch.Tensor
    sec01: torch.Tensor
    idx10: Optional[torch.Tensor] = None
    sim10: Optional[torch.Tensor] = None
    sec10: Optional[torch.Tensor] = None


def match_nn(a, b, *, normalize: bool = True, mutual: bool = False, algo: str = "auto", device=None) -> MatchResult:
    """Inner-product top-1 (+ runner-up value) of every row of ``a`` (N, D) in ``b`` (M, D); with ``mutual`` also b -> a.
    For L2-normalised rows argmax-IP == argmin-L2, which covers the reference's three matchers."""
    ctx = get_context(device)
    dev = torch.device("cuda", ctx.device)
    a = _dev_f32(a, dev, "a")
    b = _dev_f32(b, dev, "b", cols=a.shape[1])
    n, d = a.shape
    m = b.shape[0]
    if n == 0 or m == 0:
        raise ValueError(f"Invalid shape: empty descriptor set (n={n}, m={m})")
    res = MatchResult(torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.float32, device=dev),
                      torch.empty(n, dtype=torch.float32, device=dev))
    if mutual:
        res.idx10 = torch.empty(m, dtype=torch.int32, device=dev)
        res.sim10 = torch.empty(m, dtype=torch.float32, device=dev)
        res.sec10 = torch.empty(m, dtype=torch.float32, device=dev)
    flags = (_lib.NORMALIZE if normalize else 0) | (_lib.MUTUAL if mutual else 0) | _ALGO[algo]
    ctx.bind_stream()
    _lib.check(ctx.lib.vfmreg_match_nn(ctx.handle, _ptr(a), n, _ptr(b), m, d, flags, _ptr(res.idx01), _ptr(res.sim01),
                                      _ptr(res.sec01), _ptr(res.idx10), _ptr(res.sim10), _ptr(res.sec10)), "vfmreg_match_nn")
    return res


def filter_correspondences(match: MatchResult, *, min_cos: Optional[float] = None, mutual: bool = False,
                           ratio: Optional[float] = None, device=None) -> torch.Tensor:
    """(K, 2) int32 correspondences (query, match) in query order: cosine gate / mutual / ratio (see the header)."""
    ctx = get_context(device)
    n = match.idx01.shape[0]
    dev = match.idx01.device
    corr = torch.empty((n, 2), dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    ctx.bind_stream()
    _lib.check(ctx.lib.vfmreg_filter_correspondences(
        ctx.handle, _ptr(match.idx01), _ptr(match.sim01), _ptr(match.sec01), _ptr(match.idx10 if mutual else None), n,
        float("nan") if min_cos is None else float(min_cos), float("nan") if ratio is None else float(ratio), int(mutual),
        _ptr(corr), _ptr(count)), "vfmreg_filter_correspondences")
    return corr[: int(count.item())]


def l2_distances(match: MatchResult, *, device=None) -> torch.Tensor:
    """sqrt(2 - 2 s + 1e-6) per query: the L2 distance between unit descriptors in the form the reference's brute-force
    block computes it (registration_node.py:197-198).  +inf where a query has no match."""
    ctx = get_context(device)
    n = match.idx01.shape[0]
    dist = torch.empty(n, dtype=torch.float32, device=match.idx01.device)
    ctx.bind_stream()
    _lib.check(ctx.lib.vfmreg_l2_distances(ctx.handle, _ptr(match.idx01), _ptr(match.sim01), n, _ptr(dist)), "vfmreg_l2_distances")
    return dist


def select_smallest(match: MatchResult, n_points: int, *, return_distance: bool = False, device=None):
    """The ``n_points`` nearest-neighbour pairs with the smallest descriptor distance, (K, 2) int32 in query order -- the
    reference's "keep only top N correspondences (smallest distance)" (registration_node.py:212-214, 510-518; its
    ``np.argpartition`` leaves boundary ties to chance, here the lowest query index wins)."""
    ctx = get_context(device)
    n = match.idx01.shape[0]
    dev = match.idx01.device
    corr = torch.empty((n, 2), dtype=torch.int32, device=dev)
    dist = torch.empty(n, dtype=torch.float32, device=dev) if return_distance else None
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    ctx.bind_stream()
    _lib.check(ctx.lib.vfmreg_select_smallest(ctx.handle, _ptr(match.idx01), _ptr(match.sim01), n, int(n_points), _ptr(corr), _ptr(dist),
                                             _ptr(count)), "vfmreg_select_smallest")
    k = int(count.item())
    return (corr[:k], dist[:k]) if return_distance else corr[:k]


@dataclass
class RansacResult:
    T: np.ndarray            # (4, 4) float64
    best: int
    n_inliers: int
    n_corr: int
    fitness: float
    rmse: float
    counts: torch.Tensor     # (H,) int32, -1 for degenerate samples
    sumq: torch.Tensor       # (H,) int64
    mask: torch.Tensor       # (K,) bool


def ransac_kabsch(src_xyz, tgt_xyz, corr, *, sample_idx=None, n_hyp: Optional[int] = None, thresh: float = 1e4,
                  seed: int = 42, refit: bool = False, device=None) -> RansacResult:
    """Batched 3-point-Kabsch RANSAC over a correspondence list (float64 on the device)."""
    ctx = get_context(device)
    dev = torch.device("cuda", ctx.device)

    def xyz(x, name):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x))
        if x.dim() != 2 or x.shape[1] != 3:
            raise ValueError(f"Invalid shape for {name}: {tuple(x.shape)}")
        if x.dtype not in (torch.float32, torch.float64):
            x = x.to(torch.float64)
        return x.to(dev).contiguous()

    src, tgt = xyz(src_xyz, "src_xyz"), xyz(tgt_xyz, "tgt_xyz")
    if src.dtype != tgt.dtype:
        src, tgt = src.to(torch.float64), tgt.to(torch.float64)
    if isinstance(corr, np.ndarray):
        corr = torch.from_numpy(np.ascontiguousarray(corr.reshape(-1, 2), dtype=np.int32))
    corr = corr.to(device=dev, dtype=torch.int32).contiguous()
    k = corr.shape[0]
    if sample_idx is not None:
        if isinstance(sample_idx, np.ndarray):
            sample_idx = torch.from_numpy(np.ascontiguousarray(sample_idx, dtype=np.int32))
        sample_idx = sample_idx.to(device=dev, dtype=torch.int32).contiguous()
        n_hyp = sample_idx.shape[0]
    if not n_hyp or n_hyp <= 0:
        raise ValueError("ransac_kabsch: need sample_idx or a positive n_hyp")
    count = torch.full((1,), k, dtype=torch.int32, device=dev)
    t = torch.empty(16, dtype=torch.float64, device=dev)
    counts = torch.empty(n_hyp, dtype=torch.int32, device=dev)
    sumq = torch.empty(n_hyp, dtype=torch.int64, device=dev)
    mask = torch.zeros(max(k, 1), dtype=torch.uint8, device=dev)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    corr_arg = corr if k > 0 else torch.zeros((1, 2), dtype=torch.int32, device=dev)
    ctx.bind_stream()
    _lib.check(ctx.lib.vfmreg_ransac(ctx.handle, _ptr(src), _ptr(tgt), int(src.dtype == torch.float64), _ptr(corr_arg),
                                    _ptr(count), k, _ptr(sample_idx), n_hyp, seed & 0xFFFFFFFFFFFFFFFF, float(thresh),
                                    int(refit), _ptr(t), _ptr(counts), _ptr(sumq), _ptr(mask), _ptr(stats)), "vfmreg_ransac")
    st = stats.cpu().numpy()
    n_in = int(st[1])
    rmse = math.sqrt(float(st[2]) / 2.0 ** 40 * thresh * thresh / n_in) if n_in else 0.0
    return RansacResult(T=t.cpu().numpy().reshape(4, 4), best=int(st[0]), n_inliers=n_in, n_corr=int(st[3]),
                        fitness=(n_in / k if k else 0.0), rmse=rmse, counts=counts, sumq=sumq, mask=mask[:k].bool())


class KdTree:
    """Balanced k-d tree over a point cloud, built on the host and kept on the device (``vfmreg_kdtree_create``): exact
    nearest neighbours at any distance -- what the Open3D-style hypothesis score needs (``ransac_nn_all``)."""

    def __init__(self, xyz, *, device=None):
        self.ctx = get_context(device)
        x = xyz.detach().cpu().numpy() if isinstance(xyz, torch.Tensor) else np.asarray(xyz)
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.ndim != 2 or x.shape[1] != 3 or x.shape[0] == 0:
            raise ValueError(f"Invalid shape for xyz: {x.shape}")
        self.n = int(x.shape[0])
        h = C.c_void_p()
        _lib.check(self.ctx.lib.vfmreg_kdtree_create(self.ctx.handle, x.ctypes.data, self.n, C.byref(h)), "vfmreg_kdtree_create")
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.vfmreg_kdtree_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def nearest(self, queries, max_dist: float = 1e30):
        """(index (n,) int32 in the caller's order, -1 when nothing is closer than max_dist; squared distance (n,) float64)."""
        dev = torch.device("cuda", self.ctx.device)
        q = queries if isinstance(queries, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(queries))
        if q.dim() != 2 or q.shape[1] != 3:
            raise ValueError(f"Invalid shape for queries: {tuple(q.shape)}")
        q = q.to(device=dev, dtype=torch.float64).contiguous()
        n = q.shape[0]
        idx = torch.empty(n, dtype=torch.int32, device=dev)
        d2 = torch.empty(n, dtype=torch.float64, device=dev)
        self.ctx.bind_stream()
        _lib.check(self.ctx.lib.vfmreg_kdtree_nearest(self.ctx.handle, self.handle, _ptr(q), n, float(max_dist), _ptr(idx), _ptr(d2)),
                   "vfmreg_kdtree_nearest")
        return idx, d2


@dataclass
class RansacNNResult:
    T: np.ndarray            # (4, 4) float64
    best: int
    n_inliers: int           # source points whose nearest target point is closer than max_dist
    fitness: float           # n_inliers / number of source points
    rmse: float              # over those nearest-neighbour distances
    inliers: torch.Tensor    # (H,) int32 per hypothesis, -1 for degenerate samples
    sum_d2: torch.Tensor     # (H,) float64


def ransac_nn_all(src_xyz, tgt_xyz, corr, *, sample_idx=None, n_hyp: Optional[int] = None, max_dist: float = 1e4, seed: int = 42,
                  tree: Optional[KdTree] = None, device=None) -> RansacNNResult:
    """Correspondence RANSAC with the hypothesis score Open3D 0.18 uses (the reference's solver call,
    registration_node.py:312-327): hypotheses come from 3 sampled correspondences as in ``ransac_kabsch``, but each one is
    scored by transforming ALL of ``src_xyz`` and looking up every point's nearest neighbour in ``tgt_xyz`` -- fitness =
    share of points closer than ``max_dist``, rmse over those; best = higher fitness, then lower rmse.  With the
    reference's ``max_dist = 10000`` that is the hypothesis with the smallest scan -> map chamfer RMSE."""
    ctx = get_context(device)
    dev = torch.device("cuda", ctx.device)

    def xyz64(x, name):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x))
        if x.dim() != 2 or x.shape[1] != 3:
            raise ValueError(f"Invalid shape for {name}: {tuple(x.shape)}")
        return x.to(device=dev, dtype=torch.float64).contiguous()

    src, tgt = xyz64(src_xyz, "src_xyz"), xyz64(tgt_xyz, "tgt_xyz")
    if src.shape[0] == 0 or tgt.shape[0] == 0:
        raise ValueError("Invalid shape: empty cloud")
    if tree is None:
        tree = KdTree(tgt, device=ctx.device)
    if isinstance(corr, np.ndarray):
        corr = torch.from_numpy(np.ascontiguousarray(corr.reshape(-1, 2), dtype=np.int32))
    corr = corr.to(device=dev, dtype=torch.int32).contiguous()
    k = corr.shape[0]
    if sample_idx is not None:
        if isinstance(sample_idx, np.ndarray):
            sample_idx = torch.from_numpy(np.ascontiguousarray(sample_idx, dtype=np.int32))
        sample_idx = sample_idx.to(device=dev, dtype=torch.int32).contiguous()
        n_hyp = sample_idx.shape[0]
    if not n_hyp or n_hyp <= 0:
        raise ValueError("ransac_nn_all: need sample_idx or a positive n_hyp")
    count = torch.full((1,), k, dtype=torch.int32, device=dev)
    t = torch.empty(16, dtype=torch.float64, device=dev)
    inl = torch.empty(n_hyp, dtype=torch.int32, device=dev)
    sums = torch.empty(n_hyp, dtype=torch.float64, device=dev)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    corr_arg = corr if k > 0 else torch.zeros((1, 2), dtype=torch.int32, device=dev)
    ctx.bind_stream()
    _lib.check(ctx.lib.vfmreg_ransac_nn_all(ctx.handle, tree.handle, _ptr(src), src.shape[0], _ptr(src), _ptr(tgt), 1, _ptr(corr_arg),
                                           _ptr(count), k, _ptr(sample_idx), n_hyp, seed & 0xFFFFFFFFFFFFFFFF, float(max_dist), _ptr(t),
                                           _ptr(inl), _ptr(sums), _ptr(stats)), "vfmreg_ransac_nn_all")
    st = stats.cpu()
    n_in = int(st[1])
    s2 = float(st[2:3].view(torch.float64)[0])
    return RansacNNResult(T=t.cpu().numpy().reshape(4, 4), best=int(st[0]), n_inliers=n_in, fitness=n_in / src.shape[0],
                          rmse=math.sqrt(s2 / n_in) if n_in else 0.0, inliers=inl, sum_d2=sums)


@dataclass
class RegResult:
    T: np.ndarray            # (4, 4) float64, maps source into target
    corr: np.ndarray         # (K, 2) int32 (source index, target index)
    inlier_mask: np.ndarray  # (K,) bool, inliers of the returned hypothesis
    fitness: float
    rmse: float
    best_hyp: int
    n_inliers: int


def _params(normalize, min_cos, mutual, ratio, ransac_iters, inlier_thresh, seed, refit, algo):
    p = _lib.RegisterParams()
    p.flags = (_lib.NORMALIZE if normalize else 0) | (_lib.MUTUAL if mutual else 0) | _ALGO[algo]
    p.min_cos = float("nan") if min_cos is None else float(min_cos)
    p.ratio = float("nan") if ratio is None else float(ratio)
    p.n_hyp = int(ransac_iters)
    p.refit = int(refit)
    p.inlier_thresh = float(inlier_thresh)
    p.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return p


def register(source_pcd, target_pcd, src_feats, tgt_feats, *, normalize: bool = True, min_cos: Optional[float] = 0.8,
             mutual: bool = False, ratio: Optional[float] = None, ransac_iters: int = 50000, inlier_thresh: float = 1e4,
             sample_idx=None, seed: int = 42, refit: bool = False, algo: str = "auto", device=None) -> RegResult:
    """Descriptor match -> gate -> batched RANSAC, returning the 4x4 transform that maps ``source_pcd`` into
    ``target_pcd``.  Defaults are the reference's literals (cosine gate 0.8, 50000 iterations, tau = 10000,
    registration_node.py:323-326,418).  K < 3 correspondences -> identity transform, fitness 0.

    NumPy inputs run through ``vfmreg_register_host`` (copies inside the call); CUDA tensors are used in place."""
    ctx = get_context(device)
    lib = ctx.lib
    on_host = all(isinstance(x, np.ndarray) or (isinstance(x, torch.Tensor) and not x.is_cuda)
                  for x in (source_pcd, target_pcd, src_feats, tgt_feats))
    p = _params(normalize, min_cos, mutual, ratio, ransac_iters, inlier_thresh, seed, refit, algo)
    res = _lib.RegisterResult()
    if on_host:
        def h(x, name, cols=None):
            x = x.numpy() if isinstance(x, torch.Tensor) else x
            x = np.ascontiguousarray(x, dtype=np.float32)
            if x.ndim != 2 or (cols is not None and x.shape[1] != cols):
                raise ValueError(f"Invalid shape for {name}: {x.shape}")
            return x
        sx, tx = h(source_pcd, "source_pcd", 3), h(target_pcd, "target_pcd", 3)
        sf = h(src_feats, "src_feats")
        tf = h(tgt_feats, "tgt_feats", sf.shape[1])
        n, m, d = sx.shape[0], tx.shape[0], sf.shape[1]
        if sf.shape[0] != n or tf.shape[0] != m:
            raise ValueError(f"Invalid shape: {n} points vs {sf.shape[0]} descriptors / {m} vs {tf.shape[0]}")
        si = None
        if sample_idx is not None:
            si = np.ascontiguousarray(sample_idx, dtype=np.int32)
            p.n_hyp = si.shape[0]
        corr = np.empty((n, 2), dtype=np.int32)
        mask = np.empty(n, dtype=np.uint8)
        ctx.bind_stream()
        _lib.check(lib.vfmreg_register_host(ctx.handle, sx.ctypes.data, tx.ctypes.data, sf.ctypes.data, tf.ctypes.data, n, m,
                                           d, C.byref(p), si.ctypes.data if si is not None else None, corr.ctypes.data,
                                           mask.ctypes.data, C.byref(res)), "vfmreg_register_host")
        k = int(res.n_corr)
        corr_np, mask_np = corr[:k].copy(), mask[:k].astype(bool)
    else:
        dev = torch.device("cuda", ctx.device)
        sx, tx = _dev_f32(source_pcd, dev, "source_pcd", 3), _dev_f32(target_pcd, dev, "target_pcd", 3)
        sf = _dev_f32(src_feats, dev, "src_feats")
        tf = _dev_f32(tgt_feats, dev, "tgt_feats", sf.shape[1])
        n, m, d = sx.shape[0], tx.shape[0], sf.shape[1]
        if sf.shape[0] != n or tf.shape[0] != m:
            raise ValueError(f"Invalid shape: {n} points vs {sf.shape[0]} descriptors / {m} vs {tf.shape[0]}")
        si = None
        if sample_idx is not None:
            if isinstance(sample_idx, np.ndarray):
                sample_idx = torch.from_numpy(np.ascontiguousarray(sample_idx, dtype=np.int32))
            si = sample_idx.to(device=dev, dtype=torch.int32).contiguous()
            p.n_hyp = si.shape[0]
        corr = torch.empty((n, 2), dtype=torch.int32, device=dev)
        mask = torch.empty(n, dtype=torch.uint8, device=dev)
        ctx.bind_stream()
        _lib.check(lib.vfmreg_register(ctx.handle, _ptr(sx), _ptr(tx), _ptr(sf), _ptr(tf), n, m, d, C.byref(p), _ptr(si),
                                      _ptr(corr), _ptr(mask), C.byref(res)), "vfmreg_register")
        k = int(res.n_corr)
        corr_np, mask_np = corr[:k].cpu().numpy(), mask[:k].cpu().numpy().astype(bool)
    return RegResult(T=np.array(res.T, dtype=np.float64).reshape(4, 4), corr=corr_np, inlier_mask=mask_np,
                     fitness=float(res.fitness), rmse=float(res.rmse), best_hyp=int(res.best_hyp),
                     n_inliers=int(res.n_inliers))


def _register_batch_device(ctx, pairs, p, resident=None):
    """CUDA-tensor inputs: everything is enqueued back to back, one host synchronisation per batch.
    ``resident``: every pair is (source_pcd, src_feats) against that ResidentMap."""
    dev = torch.device("cuda", ctx.device)
    k = len(pairs)
    keep, ns, ms = [], [], []
    d = resident.d if resident is not None else None
    conv = {}   # pairs that share a target object share its device copy: the library prepares that map once

    def shared(x, name, cols):
        key = id(x)
        if key not in conv:
            conv[key] = (x, _dev_f32(x, dev, name, cols))
        return conv[key][1]

    for pr in pairs:
        if resident is not None:
            sx, sf = _dev_f32(pr[0], dev, "source_pcd", 3), _dev_f32(pr[1], dev, "src_feats", d)
            tx = tf = None
        else:
            sx, tx, sf = _dev_f32(pr[0], dev, "source_pcd", 3), shared(pr[1], "target_pcd", 3), _dev_f32(pr[2], dev, "src_feats")
            d = sf.shape[1] if d is None else d
            tf = shared(pr[3], "tgt_feats", d)
        if tuple(sf.shape) != (sx.shape[0], d) or (tf is not None and tf.shape[0] != tx.shape[0]):
            raise ValueError("Invalid shape: points / descriptors mismatch")
        keep.append((sx, tx, sf, tf))
        ns.append(sx.shape[0])
        ms.append(tx.shape[0] if tx is not None else resident.m)
    corr = [torch.empty((n, 2), dtype=torch.int32, device=dev) for n in ns]
    mask = [torch.empty(n, dtype=torch.uint8, device=dev) for n in ns]
    arr = lambda ptrs: (C.c_void_p * k)(*ptrs)  # noqa: E731
    res = (_lib.RegisterResult * k)()
    ctx.bind_stream()
    if resident is not None:
        _lib.check(ctx.lib.vfmreg_register_scans(
            ctx.handle, resident.handle, k, arr([x[0].data_ptr() for x in keep]), arr([x[2].data_ptr() for x in keep]),
            (C.c_int64 * k)(*ns), C.byref(p), None, 0, arr([t.data_ptr() for t in corr]), arr([t.data_ptr() for t in mask]), res),
            "vfmreg_register_scans")
    else:
        _lib.check(ctx.lib.vfmreg_register_batch(
            ctx.handle, k, arr([x[0].data_ptr() for x in keep]), arr([x[1].data_ptr() for x in keep]),
            arr([x[2].data_ptr() for x in keep]), arr([x[3].data_ptr() for x in keep]), (C.c_int64 * k)(*ns), (C.c_int64 * k)(*ms), d,
            C.byref(p), None, arr([t.data_ptr() for t in corr]), arr([t.data_ptr() for t in mask]), res), "vfmreg_register_batch")
    out = []
    for i in range(k):
        kc = int(res[i].n_corr)
        out.append(RegResult(T=np.array(res[i].T, dtype=np.float64).reshape(4, 4), corr=corr[i][:kc], inlier_mask=mask[i][:kc].bool(),
                             fitness=float(res[i].fitness), rmse=float(res[i].rmse), best_hyp=int(res[i].best_hyp),
                             n_inliers=int(res[i].n_inliers)))
    return out


def register_batch(pairs, *, normalize: bool = True, min_cos: Optional[float] = 0.8, mutual: bool = False,
                   ratio: Optional[float] = None, ransac_iters: int = 50000, inlier_thresh: float = 1e4, seed: int = 42,
                   refit: bool = False, algo: str = "auto", device=None):
    """``register`` over a list of (source_pcd, target_pcd, src_feats, tgt_feats).

    HOST arrays: the copy of pair i+1 is overlapped with the solve of pair i (``vfmreg_register_batch_host``); pinned inputs
    (``torch.Tensor.pin_memory().numpy()``) make the copies asynchronous.  CUDA tensors: all pairs are enqueued back to
    back with a single host synchronisation (``vfmreg_register_batch``); ``corr`` / ``inlier_mask`` then stay on the device.
    Consecutive pairs that pass the SAME target objects (``target_pcd`` and ``tgt_feats``) share one upload / preparation of
    that map -- the reference registers the 3-5 scans of a scene against one local map (registration_node.py:554-590).
    Mixed host / CUDA inputs are moved to the host side.  Returns a list of RegResult."""
    ctx = get_context(device)
    k = len(pairs)
    if k == 0:
        return []
    p = _params(normalize, min_cos, mutual, ratio, ransac_iters, inlier_thresh, seed, refit, algo)
    if all(isinstance(x, torch.Tensor) and x.is_cuda for pr in pairs for x in pr):
        return _register_batch_device(ctx, pairs, p)
    return _register_batch_host(ctx, pairs, p)


def _register_batch_host(ctx, pairs, p, resident=None):
    """Host buffers; ``resident``: every pair is (source_pcd, src_feats) against that ResidentMap."""
    k = len(pairs)

    def h(x, name, cols=None):
        x = x.cpu().numpy() if isinstance(x, torch.Tensor) else x
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or (cols is not None and x.shape[1] != cols):
            raise ValueError(f"Invalid shape for {name}: {x.shape}")
        return x

    conv = {}   # pairs that share a target object share its converted array: the library uploads / prepares that map once

    def shared(x, name, cols):
        key = id(x)
        if key not in conv:
            conv[key] = (x, h(x, name, cols))
        return conv[key][1]

    keep, ns, ms = [], [], []
    d = resident.d if resident is not None else None
    for pr in pairs:
        if resident is not None:
            sx, sf = h(pr[0], "source_pcd", 3), h(pr[1], "src_feats", d)
            tx = tf = None
        else:
            sx, tx, sf = h(pr[0], "source_pcd", 3), shared(pr[1], "target_pcd", 3), h(pr[2], "src_feats")
            d = sf.shape[1] if d is None else d
            tf = shared(pr[3], "tgt_feats", d)
        if sf.shape != (sx.shape[0], d) or (tf is not None and tf.shape[0] != tx.shape[0]):
            raise ValueError("Invalid shape: points / descriptors mismatch")
        keep.append((sx, tx, sf, tf))
        ns.append(sx.shape[0])
        ms.append(tx.shape[0] if tx is not None else resident.m)
    corr = [torch.empty((n, 2), dtype=torch.int32).pin_memory() for n in ns]
    mask = [torch.empty(n, dtype=torch.uint8).pin_memory() for n in ns]
    arr = lambda ptrs: (C.c_void_p * k)(*ptrs)  # noqa: E731
    res = (_lib.RegisterResult * k)()
    ctx.bind_stream()
    if resident is not None:
        _lib.check(ctx.lib.vfmreg_register_scans(
            ctx.handle, resident.handle, k, arr([x[0].ctypes.data for x in keep]), arr([x[2].ctypes.data for x in keep]),
            (C.c_int64 * k)(*ns), C.byref(p), None, 1, arr([t.data_ptr() for t in corr]), arr([t.data_ptr() for t in mask]), res),
            "vfmreg_register_scans")
    else:
        _lib.check(ctx.lib.vfmreg_register_batch_host(
            ctx.handle, k, arr([x[0].ctypes.data for x in keep]), arr([x[1].ctypes.data for x in keep]),
            arr([x[2].ctypes.data for x in keep]), arr([x[3].ctypes.data for x in keep]), (C.c_int64 * k)(*ns), (C.c_int64 * k)(*ms), d,
            C.byref(p), None, arr([t.data_ptr() for t in corr]), arr([t.data_ptr() for t in mask]), res), "vfmreg_register_batch_host")
    out = []
    for i in range(k):
        kc = int(res[i].n_corr)
        out.append(RegResult(T=np.array(res[i].T, dtype=np.float64).reshape(4, 4), corr=corr[i][:kc].numpy().copy(),
                             inlier_mask=mask[i][:kc].numpy().astype(bool), fitness=float(res[i].fitness), rmse=float(res[i].rmse),
                             best_hyp=int(res[i].best_hyp), n_inliers=int(res[i].n_inliers)))
    return out


class ResidentMap:
    """A map kept on the device for the scans of one scene (``vfmreg_map_create``): coordinates + renormalised fp32 / fp16
    descriptors, uploaded and prepared once.  The reference builds ``local_map`` once per scene
    (registration_node.py:554-580) and registers every scan of the scene against it (:587-590).

    ``target_pcd`` (M, 3) and ``tgt_feats`` (M, D): both NumPy / host tensors, or both CUDA tensors."""

    def __init__(self, target_pcd, tgt_feats, *, normalize: bool = True, algo: str = "auto", device=None):
        self.ctx = get_context(device)
        host = [not (isinstance(x, torch.Tensor) and x.is_cuda) for x in (target_pcd, tgt_feats)]
        if host[0] != host[1]:
            raise ValueError("Invalid shape: target_pcd and tgt_feats must both be host arrays or both be CUDA tensors")
        self.flags = (_lib.NORMALIZE if normalize else 0) | _ALGO[algo]
        if host[0]:
            def h(x, name, cols=None):
                x = x.numpy() if isinstance(x, torch.Tensor) else x
                x = np.ascontiguousarray(x, dtype=np.float32)
                if x.ndim != 2 or (cols is not None and x.shape[1] != cols):
                    raise ValueError(f"Invalid shape for {name}: {x.shape}")
                return x
            tx, tf = h(target_pcd, "target_pcd", 3), h(tgt_feats, "tgt_feats")
            ptrs = (tx.ctypes.data, tf.ctypes.data)
        else:
            dev = torch.device("cuda", self.ctx.device)
            tx, tf = _dev_f32(target_pcd, dev, "target_pcd", 3), _dev_f32(tgt_feats, dev, "tgt_feats")
            ptrs = (tx.data_ptr(), tf.data_ptr())
        if tf.shape[0] != tx.shape[0] or tx.shape[0] == 0:
            raise ValueError(f"Invalid shape: {tx.shape[0]} points vs {tf.shape[0]} descriptors")
        self.m, self.d = int(tx.shape[0]), int(tf.shape[1])
        h_ = C.c_void_p()
        self.ctx.bind_stream()
        _lib.check(self.ctx.lib.vfmreg_map_create(self.ctx.handle, C.c_void_p(ptrs[0]), C.c_void_p(ptrs[1]), self.m, self.d,
                                                 self.flags, int(host[0]), C.byref(h_)), "vfmreg_map_create")
        self.handle = h_

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.vfmreg_map_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return self.m

    def match(self, queries, *, min_cos: Optional[float] = None, second: bool = True) -> MatchResult:
        """Top-1 (+ runner-up value when ``second``) of every query row in the resident map.  With ``second=False`` a
        ``min_cos`` gate may be handed to the search: queries that cannot reach it report index -1, similarity -inf."""
        dev = torch.device("cuda", self.ctx.device)
        q = _dev_f32(queries, dev, "queries", self.d)
        n = q.shape[0]
        if n == 0:
            raise ValueError("Invalid shape: empty query set")
        res = MatchResult(torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.float32, device=dev),
                          torch.empty(n, dtype=torch.float32, device=dev) if second else None)
        self.ctx.bind_stream()
        _lib.check(self.ctx.lib.vfmreg_map_match(self.ctx.handle, self.handle, _ptr(q), n,
                                                float("nan") if (min_cos is None or second) else float(min_cos),
                                                _ptr(res.idx01), _ptr(res.sim01), _ptr(res.sec01)), "vfmreg_map_match")
        return res


def register_scans(resident: ResidentMap, scans, *, min_cos: Optional[float] = 0.8, mutual: bool = False,
                   ratio: Optional[float] = None, ransac_iters: int = 50000, inlier_thresh: float = 1e4, seed: int = 42,
                   refit: bool = False):
    """``register`` of every (source_pcd, src_feats) in ``scans`` against one ResidentMap (``vfmreg_register_scans``):
    host arrays are uploaded on a copy stream while the previous scan is solved; CUDA tensors are enqueued back to back.
    Returns a list of RegResult."""
    if len(scans) == 0:
        return []
    p = _params(bool(resident.flags & _lib.NORMALIZE), min_cos, mutual, ratio, ransac_iters, inlier_thresh, seed, refit, "auto")
    p.flags = (p.flags & ~0xF00) | (resident.flags & 0xF00)
    if all(isinstance(x, torch.Tensor) and x.is_cuda for pr in scans for x in pr):
        return _register_batch_device(resident.ctx, scans, p, resident)
    return _register_batch_host(resident.ctx, scans, p, resident)


# ---------------------------------------------------------------------------------------------------------------------
# 8f row 4: the TEASER++-style solve of registration_node.py:91-131
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class TeaserResult:
    T: np.ndarray              # (4, 4) float64
    clique: np.ndarray         # correspondence indices of the maximum clique, ascending
    exact: bool                # the clique search ran to completion
    iterations: int            # GNC-TLS iterations
    rotation_inliers: int
    translation_inliers: int


def teaser_solve(src_xyz, tgt_xyz, *, noise_bound: float = 0.2, cbar2: float = 1.0, gnc_factor: float = 1.4,
                 max_iterations: int = 10000, cost_threshold: float = 1e-16, max_clique_nodes: int = 0, device=None) -> TeaserResult:
    """``teaserpp_python.RobustRegistrationSolver(params).solve(src, tgt)`` with the reference's parameters
    (registration_node.py:109-121): src_xyz / tgt_xyz are the (K, 3) points of K putative correspondences.  Compatibility
    graph on the GPU, exact maximum clique, GNC-TLS rotation, TLS translation (csrc/teaser.cu; "parity unpinned")."""
    ctx = get_context(device)
    s = np.ascontiguousarray(np.asarray(src_xyz.detach().cpu() if isinstance(src_xyz, torch.Tensor) else src_xyz, dtype=np.float64))
    t = np.ascontiguousarray(np.asarray(tgt_xyz.detach().cpu() if isinstance(tgt_xyz, torch.Tensor) else tgt_xyz, dtype=np.float64))
    if s.ndim != 2 or s.shape[1] != 3 or s.shape != t.shape:
        raise ValueError(f"Invalid shape: {s.shape} / {t.shape} (two (K, 3) arrays expected)")
    k = s.shape[0]
    p = _lib.TeaserParams(float(noise_bound), float(cbar2), float(gnc_factor), float(cost_threshold), int(max_iterations), 0, int(max_clique_nodes))
    T = np.zeros(16, dtype=np.float64)
    clique = np.zeros(max(k, 1), dtype=np.int32)
    stats = np.zeros(5, dtype=np.int32)
    ctx.bind_stream()
    _lib.check(ctx.lib.vfmreg_teaser_solve(ctx.handle, s.ctypes.data, t.ctypes.data, k, C.byref(p), T.ctypes.data, clique.ctypes.data,
                                          stats.ctypes.data), "vfmreg_teaser_solve")
    return TeaserResult(T=T.reshape(4, 4), clique=clique[:stats[0]].astype(np.int64), exact=bool(stats[1]), iterations=int(stats[2]),
                        rotation_inliers=int(stats[3]), translation_inliers=int(stats[4]))


# ---------------------------------------------------------------------------------------------------------------------
# a3/a4/a5: projection + feature gather
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class CameraSpec:
    """One camera of ``project_gather`` / ``extract_features``.

    P (3, 4): pixel_h = P @ [x y z 1] in full-resolution pixels (K @ T_cam_from_lidar[:3]);
    img_hw: size of the (sub-sampled, cropped) image the black-pixel test and the feature map refer to;
    crop = (y0, x0, h, w) in sub-sampled pixels (default: whole image at the origin); see include/vfmreg_b200.h."""
    P: np.ndarray
    img_hw: tuple
    grid_hw: tuple
    crop: Optional[tuple] = None
    subsample: float = 1.0
    z_inclusive: bool = False
    float_bounds: bool = False
    black_mode: int = 1
    rot90: bool = False


def project_gather(points, cams, tokens, images=None, *, device=None):
    """Fused point->pixel projection, token-grid bilinear sampling and first-camera-wins scatter.

    points (N, 3) f32; cams: list of CameraSpec; tokens: list of (grid_h, grid_w, D) f32 tensors (one per camera);
    images: optional list of (H, W, 3) uint8 arrays (stored orientation) for the black-pixel test.
    Returns (desc (N, D) f32 cuda, cam_of_point (N,) int32, uv (N, 2) int32)."""
    ctx = get_context(device)
    dev = torch.device("cuda", ctx.device)
    pts = _dev_f32(points, dev, "points", 3)
    n = pts.shape[0]
    toks = [(torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32)) if isinstance(t, np.ndarray) else t)
            .to(device=dev, dtype=torch.float32).contiguous() for t in tokens]
    if len(toks) != len(cams) or not cams:
        raise ValueError("Invalid shape: one token grid per camera expected")
    d = toks[0].shape[-1]
    sizes = [((t.numel() + 3) // 4) * 4 for t in toks]
    tok_flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
    tok_off, o = [], 0
    for t, sz in zip(toks, sizes):
        if t.dim() != 3 or t.shape[-1] != d:
            raise ValueError(f"Invalid shape for tokens: {tuple(t.shape)}")
        tok_flat[o:o + t.numel()] = t.reshape(-1)
        tok_off.append(o)
        o += sz
    img_flat, img_off = None, None
    if images is not None:
        imgs = [torch.from_numpy(np.ascontiguousarray(im, dtype=np.uint8)) if isinstance(im, np.ndarray) else im for im in images]
        img_off, o = [], 0
        for im in imgs:
            img_off.append(o)
            o += im.numel()
        img_flat = torch.cat([im.reshape(-1).to(dev) for im in imgs])
    arr = (_lib.Camera * len(cams))()
    for i, (c, t) in enumerate(zip(cams, toks)):
        p = np.asarray(c.P, dtype=np.float64).reshape(12)
        for k in range(12):
            arr[i].P[k] = p[k]
        arr[i].img_h, arr[i].img_w = int(c.img_hw[0]), int(c.img_hw[1])
        y0, x0, h, w = c.crop if c.crop is not None else (0, 0, c.img_hw[0], c.img_hw[1])
        arr[i].crop_y0, arr[i].crop_x0, arr[i].crop_h, arr[i].crop_w = int(y0), int(x0), int(h), int(w)
        if tuple(t.shape[:2]) != tuple(c.grid_hw):
            raise ValueError(f"Invalid shape: token grid {tuple(t.shape[:2])} vs camera grid {tuple(c.grid_hw)}")
        arr[i].grid_h, arr[i].grid_w = int(c.grid_hw[0]), int(c.grid_hw[1])
        arr[i].subsample = float(c.subsample)
        arr[i].z_inclusive, arr[i].float_bounds = int(c.z_inclusive), int(c.float_bounds)
        arr[i].black_mode = int(c.black_mode) if images is not None else 0
        arr[i].rot90 = int(c.rot90)
    desc = torch.empty((n, d), dtype=torch.float32, device=dev)
    cam_of = torch.empty(n, dtype=torch.int32, device=dev)
    uv = torch.empty((n, 2), dtype=torch.int32, device=dev)
    toff = (C.c_int64 * len(cams))(*tok_off)
    ioff = (C.c_int64 * len(cams))(*img_off) if img_off is not None else None
    ctx.bind_stream()
    _lib.check(ctx.lib.vfmreg_project_gather(ctx.handle, _ptr(pts), n, arr, len(cams), _ptr(tok_flat), toff, _ptr(img_flat), ioff,
                                            d, _ptr(desc), _ptr(cam_of), _ptr(uv)), "vfmreg_project_gather")
    return desc, cam_of, uv
