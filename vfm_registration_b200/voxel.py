"""Host side of the voxel operations either side of the match-and-solve path, mirroring the
reference's python wrappers:

  voxel_down_sample(points, voxel_size)          kiss_icp/voxelization.py:27-40  -> Preprocessing.cpp:50-137
  VoxelMap (build / points / nearest)            kiss_icp/mapping.py:38-118      -> VoxelHashMap.cpp:76-168, 735-771
  register_frame(points, voxel_map, T0, d, k)    kiss_icp/registration.py:27-71  -> Registration.cpp:145-195

All three run on libvfmreg_b200.so (csrc/voxel.cu); there is no CPU fallback.  Differences from the reference, both
forced by its unspecified hash-map iteration order: down-sampled rows and ``VoxelMap.points()`` come back in input order
(the reference returns the same SET of points in tsl::robin_map order)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from .api import _ptr, get_context

MAX_NUM_ITERATIONS = 1000   # Registration.cpp:92
VOXEL_RANGE = float(1 << 20)   # the 3 x 21-bit voxel key of csrc/voxel.cu


def _usable_rows(t: torch.Tensor, voxel_size: float) -> Optional[torch.Tensor]:
    """Raw LiDAR returns contain NaN / inf rows, and a point 2^20 voxels from the origin has no key: the reference keeps
    such rows alive as undefined behaviour (`cast<int>` of a non-finite double); the library rejects the whole call.  Here
    they are dropped (with a warning) before the call.  Returns the kept row indices, or None when every row is usable."""
    xyz = t[:, :3]
    ok = torch.isfinite(xyz).all(dim=1) & ((xyz.abs().to(torch.float64) / float(voxel_size)) < VOXEL_RANGE - 1).all(dim=1)
    if bool(ok.all()):
        return None
    import warnings
    warnings.warn(f"{int((~ok).sum())} of {t.shape[0]} points are not finite or outside +-2^20 voxels: dropped", RuntimeWarning, stacklevel=3)
    return torch.nonzero(ok).squeeze(1)


def voxel_down_sample(points, voxel_size: float, *, return_index: bool = False, device=None):
    """First point (lowest row index) of every voxel, voxel = trunc((double)xyz / voxel_size).  ``points``: (N, >=3)
    float32 / float64 NumPy array or CUDA tensor; extra columns (descriptors) ride along.  Returns the kept rows in input
    order (same container type as the input), optionally with their row indices."""
    ctx = get_context(device)
    dev = torch.device("cuda", ctx.device)
    is_np = not isinstance(points, torch.Tensor)
    t = torch.as_tensor(points)
    if t.ndim != 2 or t.shape[1] < 3:
        raise ValueError("Invalid shape")   # voxelization.py:38
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    t = t.to(dev).contiguous()
    rows = _usable_rows(t, voxel_size)   # indices into the caller's array, or None
    if rows is not None:
        t = t[rows].contiguous()
    n, cols = t.shape
    if n == 0:
        out = t.cpu().numpy() if is_np else t
        return (out, np.zeros(0, dtype=np.int64)) if return_index else out
    keep = torch.empty(n, dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    ctx.bind_stream()
    _lib.check(ctx.lib.vfmreg_voxel_downsample(ctx.handle, _ptr(t), n, cols, t.element_size(), float(voxel_size), _ptr(keep),
                                              _ptr(count)), "vfmreg_voxel_downsample")
    k = int(count.item())
    out = torch.empty((k, cols), dtype=t.dtype, device=dev)
    if k:
        _lib.check(ctx.lib.vfmreg_gather_rows(ctx.handle, _ptr(t), cols * t.element_size(), _ptr(keep), _ptr(count), k, _ptr(out)),
                   "vfmreg_gather_rows")
    idx = keep[:k]
    if rows is not None:
        idx = rows[idx.long()].to(torch.int32)   # row numbers of the caller's array
    if is_np:
        out = out.cpu().numpy()
        idx = idx.cpu().numpy().astype(np.int64)
    return (out, idx) if return_index else out


class VoxelMap:
    """Device-resident voxel hash map: every voxel keeps its first ``max_points_per_voxel`` points (VoxelHashMap.cpp:735-771)."""

    def __init__(self, voxel_size: float, max_points_per_voxel: int = 20, device=None):
        self.ctx = get_context(device)
        self.voxel_size, self.max_points_per_voxel = float(voxel_size), int(max_points_per_voxel)
        h = C.c_void_p()
        _lib.check(self.ctx.lib.vfmreg_voxel_map_create(self.ctx.handle, self.voxel_size, self.max_points_per_voxel, C.byref(h)),
                   "vfmreg_voxel_map_create")
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.vfmreg_voxel_map_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _f64(self, xyz) -> torch.Tensor:
        t = torch.as_tensor(xyz)
        if t.ndim != 2 or t.shape[1] != 3:
            raise ValueError("Invalid shape")
        return t.to(device=torch.device("cuda", self.ctx.device), dtype=torch.float64).contiguous()

    def build(self, xyz) -> None:
        """Replace the content by the thinned points of ``xyz`` (N, 3)."""
        t = self._f64(xyz)
        self._rows = _usable_rows(t, self.voxel_size)   # non-finite / out-of-range rows are dropped; src_idx stays the caller's
        if self._rows is not None:
            t = t[self._rows].contiguous()
        self.ctx.bind_stream()
        _lib.check(self.ctx.lib.vfmreg_voxel_map_build(self.ctx.handle, self.handle, _ptr(t), t.shape[0]), "vfmreg_voxel_map_build")

    def __len__(self) -> int:
        return int(self.ctx.lib.vfmreg_voxel_map_size(self.handle))

    def points(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(xyz (K, 3) float64, src_idx (K,) int32 = row of the array given to build()), sorted by src_idx."""
        k = len(self)
        dev = torch.device("cuda", self.ctx.device)
        xyz = torch.empty((k, 3), dtype=torch.float64, device=dev)
        idx = torch.empty(k, dtype=torch.int32, device=dev)
        if k:
            self.ctx.bind_stream()
            _lib.check(self.ctx.lib.vfmreg_voxel_map_points(self.ctx.handle, self.handle, _ptr(xyz), _ptr(idx)), "vfmreg_voxel_map_points")
            order = torch.argsort(idx)
            xyz, idx = xyz[order], idx[order]
            if getattr(self, "_rows", None) is not None:
                idx = self._rows[idx.long()].to(torch.int32)
        return xyz, idx

    def nearest(self, query, max_dist: float):
        """GetClosestNeighbor for every query row: (target xyz (n, 3), valid mask (n,), squared distance (n,))."""
        q = self._f64(query)
        n = q.shape[0]
        dev = q.device
        nn = torch.full((n,), -1, dtype=torch.int32, device=dev)
        d2 = torch.full((n,), -1.0, dtype=torch.float64, device=dev)
        if n and len(self):
            self.ctx.bind_stream()
            _lib.check(self.ctx.lib.vfmreg_voxel_map_nearest(self.ctx.handle, self.handle, _ptr(q), n, float(max_dist), _ptr(nn), _ptr(d2)),
                       "vfmreg_voxel_map_nearest")
        raw_xyz = torch.empty((len(self), 3), dtype=torch.float64, device=dev)
        if len(self):
            _lib.check(self.ctx.lib.vfmreg_voxel_map_points(self.ctx.handle, self.handle, _ptr(raw_xyz), None), "vfmreg_voxel_map_points")
        valid = nn >= 0
        tgt = raw_xyz[nn.clamp(min=0).long()] if len(self) else torch.zeros((n, 3), dtype=torch.float64, device=dev)
        return tgt, valid, d2


def register_frame(points, voxel_map: "VoxelMap", initial_guess, max_correspondance_distance: float, kernel: float, *,
                   max_iterations: int = MAX_NUM_ITERATIONS, return_info: bool = False):
    """Point-to-point ICP of ``points`` (N, 3) against the voxel map starting at ``initial_guess`` (4x4); returns the
    refined 4x4 float64 pose (T_icp @ initial_guess), as kiss_icp.registration.register_frame does for (N, 3) input."""
    pts = torch.as_tensor(points)
    if pts.ndim != 2 or pts.shape[1] != 3:
        raise ValueError("Invalid shape")   # registration.py:37 (descriptor-carrying frames are the VFM-ICP variant)
    T0 = np.ascontiguousarray(np.asarray(initial_guess, dtype=np.float64))
    if T0.shape != (4, 4):
        raise ValueError("Invalid shape")
    ctx = voxel_map.ctx
    pts = pts.to(device=torch.device("cuda", ctx.device), dtype=torch.float64).contiguous()
    T = np.zeros((4, 4), dtype=np.float64)
    iters, corr = C.c_int32(0), C.c_int32(0)
    ctx.bind_stream()
    _lib.check(ctx.lib.vfmreg_register_frame(ctx.handle, voxel_map.handle, _ptr(pts), pts.shape[0], T0.ctypes.data_as(C.c_void_p),
                                            float(max_correspondance_distance), float(kernel), int(max_iterations),
                                            T.ctypes.data_as(C.c_void_p), C.byref(iters), C.byref(corr)), "vfmreg_register_frame")
    if return_info:
        return T, {"iterations": int(iters.value), "correspondences": int(corr.value)}
    return T


def register_frame_vfm(points_xyz, voxel_map: "VoxelMap", vfm_src, vfm_tgt, initial_guess, max_correspondance_distance: float,
                       kernel: float, *, max_iterations: int = MAX_NUM_ITERATIONS, return_info: bool = False):
    """The descriptor-carrying RegisterFrame ("VFM-ICP", Registration.cpp:197-382) given its descriptor correspondences:
    ``vfm_src`` (K, 3) source points before the initial guess is applied, ``vfm_tgt`` (K, 3) matched map points.  Loop 1
    refines on those (MAD pruning), loop 2 is the vanilla ICP of ``register_frame`` with the remaining iteration budget."""
    ctx = voxel_map.ctx
    dev = torch.device("cuda", ctx.device)

    def f64(x):
        t = torch.as_tensor(x)
        if t.ndim != 2 or t.shape[1] != 3:
            raise ValueError("Invalid shape")
        return t.to(device=dev, dtype=torch.float64).contiguous()
    pts, s, t = f64(points_xyz), f64(np.zeros((0, 3)) if vfm_src is None else vfm_src), f64(np.zeros((0, 3)) if vfm_tgt is None else vfm_tgt)
    if s.shape != t.shape:
        raise ValueError("Invalid shape")
    T0 = np.ascontiguousarray(np.asarray(initial_guess, dtype=np.float64))
    if T0.shape != (4, 4):
        raise ValueError("Invalid shape")
    T = np.zeros((4, 4), dtype=np.float64)
    it1, it2, kept = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    ctx.bind_stream()
    _lib.check(ctx.lib.vfmreg_register_frame_vfm(ctx.handle, voxel_map.handle, _ptr(pts), pts.shape[0], _ptr(s), _ptr(t), s.shape[0],
                                                T0.ctypes.data_as(C.c_void_p), float(max_correspondance_distance), float(kernel),
                                                int(max_iterations), T.ctypes.data_as(C.c_void_p), C.byref(it1), C.byref(it2),
                                                C.byref(kept)), "vfmreg_register_frame_vfm")
    if return_info:
        return T, {"vfm_iterations": int(it1.value), "iterations": int(it2.value), "vfm_kept": int(kept.value)}
    return T
