"""Compatibility shim: the reference's own entry points for the hot path, same names / array layouts / error behaviour,
running on libvfmreg_b200.so.  A user of src/vfm-reg can import these instead of the kiss_icp / Open3D ones.

  VoxelHashMap.get_vfm_correspondences   kiss_icp/mapping.py:120-131 -> pybind :128 -> VoxelHashMap.cpp:461-626
  RegistrationNode.ransac_registration   src/vfm-reg/src/registration_node.py:273-357 (method='vfm')
  RegistrationNode.compute_vfm_correspondences / compute_errors / compute_success_rate   :396-425, :997-1025

Points and descriptors travel as ONE (N, 3 + D) array, as in the reference (registration_node.py:579).

Not reproduced here ("next" rows): the voxel down-sampling steps around the matcher
(registration_node.py:399-403,414 -- callers pass clouds at the density they want matched), first-come voxel thinning in
``add_points`` (VoxelHashMap.cpp:746-757) and the ICP refinement (``run_icp=True`` raises NotImplementedError)."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import api, metrics


class VoxelHashMap:
    """Holds the map cloud with its descriptors on the device, like the reference object holds it in C++."""

    def __init__(self, voxel_size: float = 1.0, max_distance: float = 100.0, max_points_per_voxel: int = 20, device=None):
        self.voxel_size, self.max_distance, self.max_points_per_voxel = voxel_size, max_distance, max_points_per_voxel
        self._ctx = api.get_context(device)
        self._xyz = np.zeros((0, 3), dtype=np.float64)
        self._feat: Optional[torch.Tensor] = None

    def add_points(self, points: np.ndarray) -> None:
        points = np.asarray(points)
        if points.ndim != 2 or points.shape[1] < 3:
            raise ValueError("Invalid shape")  # mapping.py:84-85
        self._xyz = np.concatenate([self._xyz, points[:, :3].astype(np.float64)], axis=0)
        if points.shape[1] > 3:
            f = torch.from_numpy(np.ascontiguousarray(points[:, 3:], dtype=np.float32)).to(f"cuda:{self._ctx.device}")
            self._feat = f if self._feat is None else torch.cat([self._feat, f], dim=0)

    def empty(self) -> bool:
        return self._xyz.shape[0] == 0

    def point_cloud(self) -> np.ndarray:
        return self._xyz.copy()

    def get_vfm_correspondences(self, points: np.ndarray, max_correspondance_distance: float) -> Tuple[np.ndarray, np.ndarray]:
        """(src_xyz[K,3] f64, tgt_xyz[K,3] f64) of the queries whose top-1 cosine is >= the threshold, in query order.
        The parameter keeps the reference's (mis)name: it is the minimum cosine similarity (mapping.py:120-131)."""
        points = np.asarray(points)
        if points.ndim != 2 or self._feat is None or points.shape[1] != 3 + self._feat.shape[1]:
            raise ValueError("Invalid shape")
        if points.shape[0] == 0 or self.empty():
            return np.zeros((0, 3)), np.zeros((0, 3))  # the reference has UB here (VoxelHashMap.cpp:464)
        m = api.match_nn(points[:, 3:], self._feat, normalize=True, device=self._ctx.device)
        corr = api.filter_correspondences(m, min_cos=float(max_correspondance_distance), device=self._ctx.device).cpu().numpy()
        return points[corr[:, 0], :3].astype(np.float64), self._xyz[corr[:, 1]]


class RegistrationNode:
    """The two hot methods of the reference's experiment driver plus its error bookkeeping."""

    def __init__(self, ransac_iters: int = 50000, max_correspondence_distance: float = 10000.0, min_cosine: float = 0.8,
                 seed: int = 42, device=None):
        self.ransac_iters, self.max_dist, self.min_cosine, self.seed, self.device = (ransac_iters, max_correspondence_distance,
                                                                                   min_cosine, seed, device)
        self.rot_errors, self.trans_errors = {}, {}

    def compute_vfm_correspondences(self, voxel_map: np.ndarray, raw_scan: np.ndarray, initial_pose=np.eye(4)):
        vmap = VoxelHashMap(device=self.device)
        vmap.add_points(voxel_map)
        return vmap.get_vfm_correspondences(metrics.transform_pcl(raw_scan, initial_pose), self.min_cosine)

    def ransac_registration(self, voxel_map: np.ndarray, raw_scan: np.ndarray, method: str, run_icp: bool = False):
        if method != "vfm":
            if method in ("fpfh", "dip", "gedi", "fcgf", "gcl", "spinnet"):
                raise NotImplementedError(f"baseline descriptor '{method}' is outside the VFM hot path")
            raise ValueError(f"Invalid method: {method}")  # registration_node.py:284-285
        if run_icp:
            raise NotImplementedError("ICP refinement (register_frame) is a 'next' row, SURVEY.md section 8f")
        voxel_map, raw_scan = np.asarray(voxel_map), np.asarray(raw_scan)
        if voxel_map.ndim != 2 or raw_scan.ndim != 2 or voxel_map.shape[1] != raw_scan.shape[1] or raw_scan.shape[1] <= 3:
            raise ValueError("Invalid shape")
        r = api.register(raw_scan[:, :3], voxel_map[:, :3], raw_scan[:, 3:], voxel_map[:, 3:], normalize=True,
                         min_cos=self.min_cosine, ransac_iters=self.ransac_iters, inlier_thresh=self.max_dist, seed=self.seed,
                         device=self.device)
        return r.T, None

    def compute_errors(self, pose, gt_pose, method: str):
        t, r = metrics.compute_errors(np.asarray(pose), np.asarray(gt_pose))
        self.rot_errors.setdefault(method, []).append(r)
        self.trans_errors.setdefault(method, []).append(t)
        return t, r

    def compute_success_rate(self, method: str, translation_threshold, rotation_threshold) -> float:
        return metrics.success_rate(self.trans_errors[method], self.rot_errors[method], translation_threshold, rotation_threshold)
