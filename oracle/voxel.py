"""ORACLE (test infrastructure only -- never imported by the product): CPU restatement of the voxel operations either
side of the hot path, all citations relative to /root/reference/src/kiss-icp/cpp/kiss_icp/core.

  voxel_down_sample      Preprocessing.cpp:50-137   first point per voxel, voxel = (p / voxel_size).cast<int>()
  VoxelHashMapOracle     VoxelHashMap.cpp:735-771   AddPoints + VoxelBlock::AddPoint (VoxelHashMap.hpp:45-52): first
                                                    max_points_per_voxel points per voxel, insertion order
    .closest_neighbor    VoxelHashMap.cpp:79-136    27 voxels in (i, j, k) ascending order, strict '<'
    .get_correspondences VoxelHashMap.cpp:137-166   keep iff (closest - point).norm() < max distance
  register_frame         Registration.cpp:96-195    BuildLinearSystem + ldlt solve + SE3::exp, |dx| < 1e-4, <= 1000 iterations

PARITY UNPINNED: the reference's C++ (Eigen, Sophus, TBB, tsl::robin_map -- none of them on this machine, its CMake
downloads them) cannot be built offline and ships no golden vectors, so this restatement is checked against properties
(brute-force nearest neighbours, known SE(3) recovery, scipy's matrix exponential) instead of reference outputs.
Orders that the reference leaves to tsl::robin_map / tbb::parallel_reduce are fixed here: outputs by input index, sums in
index order."""
from __future__ import annotations

import numpy as np


def voxel_index(xyz: np.ndarray, voxel_size: float) -> np.ndarray:
    """(p / voxel_size).cast<int>(): float64 division, truncation toward zero (Preprocessing.cpp:58)."""
    return np.trunc(np.asarray(xyz, dtype=np.float64)[:, :3] / float(voxel_size)).astype(np.int64)


def voxel_down_sample(points: np.ndarray, voxel_size: float, return_index: bool = False):
    """Rows of the first point of every voxel, in input order (the reference: same set, robin_map order)."""
    points = np.asarray(points)
    seen, keep = set(), []
    for i, v in enumerate(map(tuple, voxel_index(points, voxel_size))):
        if v not in seen:          # `if (grid.contains(voxel)) continue;`
            seen.add(v)
            keep.append(i)
    keep = np.asarray(keep, dtype=np.int64)
    return (points[keep], keep) if return_index else points[keep]


class VoxelHashMapOracle:
    def __init__(self, voxel_size: float, max_points_per_voxel: int = 20):
        self.voxel_size, self.max_points = float(voxel_size), int(max_points_per_voxel)
        self.map = {}      # voxel -> list of (xyz float64[3], insertion id)
        self._next = 0

    def add_points(self, xyz: np.ndarray) -> None:
        xyz = np.asarray(xyz, dtype=np.float64)[:, :3]
        for p, v in zip(xyz, map(tuple, voxel_index(xyz, self.voxel_size))):
            block = self.map.setdefault(v, [])
            if len(block) < self.max_points:   # VoxelBlock::AddPoint
                block.append((p.copy(), self._next))
            self._next += 1

    def point_cloud(self):
        """(xyz, insertion id) of the kept points sorted by insertion id."""
        items = sorted((pid, p) for block in self.map.values() for p, pid in block)
        if not items:
            return np.zeros((0, 3)), np.zeros(0, dtype=np.int64)
        return np.stack([p for _, p in items]), np.asarray([i for i, _ in items], dtype=np.int64)

    def closest_neighbor(self, point: np.ndarray):
        kx, ky, kz = (int(c) for c in np.trunc(point / self.voxel_size))   # static_cast<int>
        best, best_d2 = None, np.finfo(np.float64).max
        for i in range(kx - 1, kx + 2):
            for j in range(ky - 1, ky + 2):
                for k in range(kz - 1, kz + 2):
                    for p, _ in self.map.get((i, j, k), ()):
                        d = p - point
                        d2 = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]
                        if d2 < best_d2:
                            best, best_d2 = p, d2
        return best, best_d2

    def get_correspondences(self, points: np.ndarray, max_dist: float):
        src, tgt = [], []
        for p in np.asarray(points, dtype=np.float64):
            q, d2 = self.closest_neighbor(p)
            if q is not None and np.sqrt(d2) < max_dist:
                src.append(p)
                tgt.append(q)
        if not src:
            return np.zeros((0, 3)), np.zeros((0, 3))
        return np.stack(src), np.stack(tgt)


def hat(w):
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


def se3_exp(dx: np.ndarray) -> np.ndarray:
    """Sophus::SE3d::exp: dx = (upsilon, omega); R = exp(hat(omega)), t = V upsilon."""
    u, w = dx[:3], dx[3:]
    th2 = float(w @ w)
    th = np.sqrt(th2)
    if th < 1e-5:
        A, B, Cc = 1.0 - th2 / 6.0, 0.5 - th2 / 24.0, 1.0 / 6.0 - th2 / 120.0
    else:
        A, B, Cc = np.sin(th) / th, (1.0 - np.cos(th)) / th2, (th - np.sin(th)) / (th2 * th)
    W = hat(w)
    T = np.eye(4)
    T[:3, :3] = np.eye(3) + A * W + B * (W @ W)
    T[:3, 3] = (np.eye(3) + B * W + Cc * (W @ W)) @ u
    return T


def build_linear_system(src: np.ndarray, tgt: np.ndarray, kernel: float):
    """BuildLinearSystem (Registration.cpp:96-140): J = [I | -hat(s)], r = s - t, w = kernel^2 / (kernel + |r|^2)^2."""
    JTJ, JTr = np.zeros((6, 6)), np.zeros(6)
    for s, t in zip(src, tgt):
        r = s - t
        J = np.hstack([np.eye(3), -hat(s)])
        w = kernel * kernel / (kernel + float(r @ r)) ** 2
        JTJ += J.T @ (w * J)
        JTr += J.T @ (w * r)
    return JTJ, JTr


def register_frame(frame: np.ndarray, vmap: VoxelHashMapOracle, initial_guess: np.ndarray, max_dist: float, kernel: float,
                   max_iterations: int = 1000, return_info: bool = False):
    T0 = np.asarray(initial_guess, dtype=np.float64)
    if not vmap.map:
        return (T0.copy(), {"iterations": 0, "correspondences": 0}) if return_info else T0.copy()
    source = np.asarray(frame, dtype=np.float64) @ T0[:3, :3].T + T0[:3, 3]
    T_icp = np.eye(4)
    iters, ncorr = 0, 0
    for _ in range(max_iterations):
        src, tgt = vmap.get_correspondences(source, max_dist)
        ncorr = len(src)
        if ncorr == 0:
            break
        JTJ, JTr = build_linear_system(src, tgt, kernel)
        dx = np.linalg.solve(JTJ, -JTr)
        est = se3_exp(dx)
        source = source @ est[:3, :3].T + est[:3, 3]
        T_icp = est @ T_icp
        iters += 1
        if np.linalg.norm(dx) < 1e-4:
            break
    T = T_icp @ T0
    return (T, {"iterations": iters, "correspondences": ncorr}) if return_info else T


def _median_nth(values: np.ndarray) -> float:
    """std::nth_element median with the even-size average of the two middle order statistics (Registration.cpp:283-293)."""
    s = np.sort(np.asarray(values, dtype=np.float64))
    n = len(s) // 2
    return float(s[n]) if len(s) & 1 else float((s[n] + s[n - 1]) / 2)


def vfm_icp_first_loop(src_raw: np.ndarray, tgt: np.ndarray, initial_guess: np.ndarray, kernel: float, max_iterations: int = 1000):
    """Registration.cpp:236-329: Gauss-Newton on the fixed descriptor correspondences with MAD pruning.  Returns
    (T_icp @ initial_guess, j, surviving (src, tgt))."""
    T0 = np.asarray(initial_guess, dtype=np.float64)
    src = np.asarray(src_raw, dtype=np.float64).reshape(-1, 3) @ T0[:3, :3].T + T0[:3, 3]
    tgt = np.asarray(tgt, dtype=np.float64).reshape(-1, 3)
    T_icp = np.eye(4)
    prev = np.linalg.norm(src - tgt, axis=1).mean() if len(src) else np.nan
    j = 0
    while j < max_iterations:
        if len(src) == 0:
            break
        JTJ, JTr = build_linear_system(src, tgt, kernel)
        est = se3_exp(np.linalg.solve(JTJ, -JTr))
        src = src @ est[:3, :3].T + est[:3, 3]
        T_icp = est @ T_icp
        d = np.linalg.norm(src - tgt, axis=1)
        mean = d.mean()
        median = _median_nth(d)
        mad = _median_nth(np.abs(d - median)) * 1.4826
        keep = np.abs(d - median) < 1.5 * mad
        src, tgt = src[keep], tgt[keep]
        if abs(prev - mean) < 0.01:   # EUCL_DIST_THRESHOLD_: break does not advance j
            break
        prev = mean
        j += 1
    return T_icp @ T0, j, (src, tgt)


def register_frame_vfm(frame_xyz: np.ndarray, vmap: VoxelHashMapOracle, vfm_src_raw: np.ndarray, vfm_tgt: np.ndarray,
                       initial_guess: np.ndarray, max_dist: float, kernel: float, max_iterations: int = 1000, return_info: bool = False):
    """RegisterFrame(VectorNd...) (Registration.cpp:197-382) given the descriptor correspondences it starts from."""
    T0 = np.asarray(initial_guess, dtype=np.float64)
    if not vmap.map:
        return (T0.copy(), {"vfm_iterations": 0, "iterations": 0, "vfm_kept": 0}) if return_info else T0.copy()
    T1, j, (s, _) = vfm_icp_first_loop(vfm_src_raw, vfm_tgt, T0, kernel, max_iterations)
    info = {"vfm_iterations": j, "vfm_kept": len(s), "iterations": 0}
    T = T1
    if max_iterations - j > 0 and len(frame_xyz):
        T, i2 = register_frame(frame_xyz, vmap, T1, max_dist, kernel, max_iterations - j, return_info=True)
        info["iterations"] = i2["iterations"]
    return (T, info) if return_info else T


def build_local_map(map_poses, map_point_clouds, voxel_size: float = 0.25, feat_dim: int = 384, split_above: int = 1_000_000) -> np.ndarray:
    """registration_node.py:557-581 with the oracle's voxel_down_sample (rows in input order)."""
    frames = []
    for pose, pcl in zip(map_poses, map_point_clouds):
        pcl = np.asarray(pcl)
        pcl = pcl[np.sum(pcl[:, 3:], axis=1) > 0]
        pcl = voxel_down_sample(pcl, voxel_size).astype(pcl.dtype) if len(pcl) else pcl
        pose = np.asarray(pose, dtype=np.float64)
        xyz = pcl[:, :3].astype(np.float64) @ pose[:3, :3].T + pose[:3, 3]      # vfm_reg/utils.py:47-54
        frames.append(np.c_[xyz, pcl[:, 3:]].astype(pcl.dtype))
    local = np.concatenate(frames, axis=0).astype(np.float32)
    if local.shape[0] > split_above:
        mean_x = np.mean(local[:, :3], axis=0)[0]
        local = np.concatenate([voxel_down_sample(local[local[:, 0] > mean_x], voxel_size),
                                voxel_down_sample(local[local[:, 0] <= mean_x], voxel_size)], axis=0)
    else:
        local = voxel_down_sample(local, voxel_size)
    return local[:, :3 + feat_dim]


# ---- vectorised restatements for full-size checks (same results as the loops above, see tests/test_oracle_voxel.py) ----
def _voxel_keys(xyz: np.ndarray, voxel_size: float) -> np.ndarray:
    v = voxel_index(xyz, voxel_size) + (1 << 20)
    return (v[:, 0] << 42) | (v[:, 1] << 21) | v[:, 2]


def voxel_down_sample_index_fast(points: np.ndarray, voxel_size: float) -> np.ndarray:
    """Indices (ascending) of the first point of every voxel."""
    _, first = np.unique(_voxel_keys(points, voxel_size), return_index=True)
    return np.sort(first)


def voxel_map_kept_index_fast(xyz: np.ndarray, voxel_size: float, max_points: int) -> np.ndarray:
    """Indices (ascending) of the points a VoxelHashMap keeps: the first max_points of every voxel in insertion order."""
    keys = _voxel_keys(xyz, voxel_size)
    order = np.argsort(keys, kind="stable")          # groups voxels, insertion order inside a voxel
    ks = keys[order]
    start = np.r_[0, np.flatnonzero(ks[1:] != ks[:-1]) + 1]
    rank = np.arange(len(ks)) - np.repeat(start, np.diff(np.r_[start, len(ks)]))
    return np.sort(order[rank < max_points])
