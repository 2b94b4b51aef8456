"""Compatibility shim: the reference's own entry points for the hot path and for the steps either side of it, same names /
array layouts / error behaviour, running on libvfmreg_b200.so.  A user of src/vfm-reg can import these instead of the
kiss_icp / Open3D ones.

  voxel_down_sample                      kiss_icp/voxelization.py:27-40 -> Preprocessing.cpp:50-137
  VoxelHashMap.add_points / point_cloud / get_correspondences / get_vfm_correspondences
                                         kiss_icp/mapping.py:38-131 -> VoxelHashMap.cpp:76-168, 461-626, 735-771
  register_frame                         kiss_icp/registration.py:27-71 -> Registration.cpp:145-195 ((N, 3) frames) and :197-382 (VFM-ICP)
  RegistrationNode.ransac_registration   src/vfm-reg/src/registration_node.py:273-357 (method='vfm', run_icp)
  RegistrationNode.compute_vfm_correspondences / compute_errors / compute_success_rate   :396-425, :997-1025

Points and descriptors travel as ONE (N, 3 + D) array, as in the reference (registration_node.py:579).

Where the reference's result depends on tsl::robin_map iteration order (the ORDER of down-sampled rows and of
``point_cloud()``), rows come back in input order here; the sets of points are the same.  The baseline descriptors
(FPFH, DIP, GeDi, FCGF, GCL, SpinNet) are not part of this build."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import api, metrics
from . import voxel as _voxel

voxel_down_sample = _voxel.voxel_down_sample


def register_frame(points, voxel_map: "VoxelHashMap", initial_guess, max_correspondance_distance: float, kernel: float):
    """kiss_icp.registration.register_frame: (N, 3) frames run the point-to-point ICP (Registration.cpp:145-195); frames
    that carry descriptors, (N, 3 + D), run the VFM-ICP overload (Registration.cpp:197-382): descriptor correspondences of
    the 5 m-voxelised, initially-posed frame at cosine >= 0.8, a correspondence-driven loop with MAD pruning, then the
    vanilla loop."""
    points = np.asarray(points)
    if points.ndim != 2 or points.shape[1] < 3:
        raise ValueError("Invalid shape")
    T0 = np.asarray(initial_guess, dtype=np.float64)
    if points.shape[1] == 3:
        if voxel_map.empty():
            return T0.copy()                                                 # "if (voxel_map.Empty()) return initial_guess" (Registration.cpp:150)
        return _voxel.register_frame(points, voxel_map._map3, T0, max_correspondance_distance, kernel)
    if voxel_map.empty_n():
        return T0.copy()                                                     # "if (voxel_map.EmptyN()) return initial_guess"
    source = metrics.transform_pcl(points, T0)
    vox, idx = voxel_down_sample(source, 5.0, return_index=True)
    if vox.shape[0] < 100:                                                   # "Voxelized too sparse. Keep input."
        vox, idx = source, np.arange(points.shape[0])
    m = voxel_map.resident().match(vox[:, 3:], min_cos=0.8, second=False)
    corr = api.filter_correspondences(m, min_cos=0.8, device=voxel_map._ctx.device).cpu().numpy()
    vfm_src = points[idx[corr[:, 0]], :3].astype(np.float64)                 # before the initial guess (applied on the device)
    vfm_tgt = voxel_map._xyzn[corr[:, 1]]
    core = voxel_map._map3 if len(voxel_map._map3) else voxel_map._mapn
    return _voxel.register_frame_vfm(points[:, :3], core, vfm_src, vfm_tgt, T0, max_correspondance_distance, kernel)


def find_correspondences(feats0, feats1, n_points: int = 5000, mutual_filter: bool = True, device=None):
    """The nested ``find_correspondences`` of compute_correspondences (registration_node.py:482-538, adapted there from
    TEASER++): nearest neighbours 0 -> 1; with ``mutual_filter`` the pairs that are also nearest neighbours 1 -> 0, else the
    ``min(n_points, len - 1)`` pairs with the smallest distance.  Returns (idx0, idx1) int64 arrays in query order.

    The search is the inner-product search of L2-renormalised rows, i.e. the reference's L2 nearest neighbour for
    unit-norm descriptors (which the learned baseline descriptors it is used with are)."""
    m = api.match_nn(feats0, feats1, normalize=True, mutual=mutual_filter, device=device)
    if mutual_filter:
        corr = api.filter_correspondences(m, mutual=True, device=device)
    else:
        corr = api.select_smallest(m, min(int(n_points), m.idx01.shape[0] - 1), device=device)
    corr = corr.cpu().numpy().astype(np.int64)
    return corr[:, 0], corr[:, 1]


class VoxelHashMap:
    """The reference object's two point stores: ``map_`` for (N, 3) clouds and ``map_n_`` for descriptor-carrying clouds,
    each keeping the first ``max_points_per_voxel`` points per voxel; both live on the device."""

    def __init__(self, voxel_size: float = 1.0, max_distance: float = 100.0, max_points_per_voxel: int = 20, device=None):
        self.voxel_size, self.max_distance, self.max_points_per_voxel = voxel_size, max_distance, max_points_per_voxel
        self._ctx = api.get_context(device)
        self._map3 = _voxel.VoxelMap(voxel_size, max_points_per_voxel, device=self._ctx.device)
        self._mapn = _voxel.VoxelMap(voxel_size, max_points_per_voxel, device=self._ctx.device)
        self._xyz3 = np.zeros((0, 3), dtype=np.float64)   # kept points of map_, insertion order
        self._xyzn = np.zeros((0, 3), dtype=np.float64)   # kept points of map_n_
        self._feat: Optional[torch.Tensor] = None          # their descriptors (device, float32)
        self._resident: Optional[api.ResidentMap] = None   # map_n_ prepared for the descriptor search (built on first use)

    def clear(self) -> None:
        self.__init__(self.voxel_size, self.max_distance, self.max_points_per_voxel, self._ctx.device)

    def add_points(self, points: np.ndarray) -> None:
        points = np.asarray(points)
        if points.ndim != 2 or points.shape[1] < 3:
            raise ValueError("Invalid shape")  # mapping.py:84-85
        if points.shape[0] == 0:
            return
        new_xyz = points[:, :3].astype(np.float64)
        if points.shape[1] == 3:
            # cumulative insertion == thinning of [kept so far, new points]: a kept point stays kept, order is preserved
            allx = np.concatenate([self._xyz3, new_xyz], axis=0)
            self._map3.build(allx)
            xyz, idx = self._map3.points()
            self._xyz3 = xyz.cpu().numpy()
            return
        k_old = self._xyzn.shape[0]
        allx = np.concatenate([self._xyzn, new_xyz], axis=0)
        self._mapn.build(allx)
        xyz, idx = self._mapn.points()
        self._xyzn = xyz.cpu().numpy()
        f_new = torch.from_numpy(np.ascontiguousarray(points[:, 3:], dtype=np.float32)).to(f"cuda:{self._ctx.device}")
        if self._feat is not None and self._feat.shape[1] != f_new.shape[1]:
            raise ValueError("Invalid shape")
        f_all = f_new if self._feat is None else torch.cat([self._feat, f_new], dim=0)
        assert f_all.shape[0] == k_old + points.shape[0]
        self._feat = f_all[idx.long()]
        self._resident = None

    def resident(self) -> "api.ResidentMap":
        """map_n_ as a device-resident, renormalised descriptor map: prepared once, then searched by every
        get_vfm_correspondences / registration call until the next add_points (the reference re-dumps and re-normalises the
        whole map inside every GetVFMCorrespondences call, VoxelHashMap.cpp:465-482)."""
        if self._feat is None:
            raise ValueError("Invalid shape")
        if self._resident is None:
            xyz = torch.from_numpy(self._xyzn.astype(np.float32)).to(self._feat.device)
            self._resident = api.ResidentMap(xyz, self._feat, device=self._ctx.device)
        return self._resident

    def empty(self) -> bool:
        return self._xyz3.shape[0] == 0

    def empty_n(self) -> bool:
        return self._xyzn.shape[0] == 0

    def point_cloud(self) -> np.ndarray:
        return self._xyz3.copy()

    def point_cloud_n(self) -> np.ndarray:
        if self._feat is None:
            return np.zeros((0, 3))
        return np.c_[self._xyzn, self._feat.cpu().numpy().astype(np.float64)]

    def get_correspondences(self, points: np.ndarray, max_correspondance_distance: float) -> Tuple[np.ndarray, np.ndarray]:
        """(source, target) of the points whose closest map point (27-voxel search) is nearer than the distance."""
        points = np.asarray(points)
        if points.ndim != 2 or points.shape[1] < 3:
            raise ValueError("Invalid shape")
        core = self._map3 if len(self._map3) else self._mapn
        if points.shape[0] == 0 or len(core) == 0:
            return np.zeros((0, 3)), np.zeros((0, 3))
        tgt, valid, _ = core.nearest(points[:, :3], max_correspondance_distance)
        valid = valid.cpu().numpy()
        return points[valid, :3].astype(np.float64), tgt.cpu().numpy()[valid]

    def get_vfm_correspondences(self, points: np.ndarray, max_correspondance_distance: float) -> Tuple[np.ndarray, np.ndarray]:
        """(src_xyz[K,3] f64, tgt_xyz[K,3] f64) of the queries whose top-1 cosine is >= the threshold, in query order.
        The parameter keeps the reference's (mis)name: it is the minimum cosine similarity (mapping.py:120-131)."""
        points = np.asarray(points)
        if points.ndim != 2 or points.shape[1] <= 3:
            raise ValueError("Invalid shape")
        if points.shape[0] == 0 or self.empty_n():
            return np.zeros((0, 3)), np.zeros((0, 3))  # the reference has UB here (VoxelHashMap.cpp:464)
        if points.shape[1] != 3 + self._feat.shape[1]:
            raise ValueError("Invalid shape")
        min_cos = float(max_correspondance_distance)
        m = self.resident().match(points[:, 3:], min_cos=min_cos, second=False)
        corr = api.filter_correspondences(m, min_cos=min_cos, device=self._ctx.device).cpu().numpy()
        return points[corr[:, 0], :3].astype(np.float64), self._xyzn[corr[:, 1]]


class RegistrationNode:
    """The two hot methods of the reference's experiment driver plus its error bookkeeping."""

    def __init__(self, ransac_iters: int = 50000, max_correspondence_distance: float = 10000.0, min_cosine: float = 0.8,
                 seed: int = 42, device=None, *, voxel_size: float = 1.0, max_points_per_voxel: int = 20, max_range: float = 100.0,
                 initial_threshold: float = 2.0, preprocess: bool = True, score: str = "auto"):
        """``voxel_size`` / ``max_points_per_voxel`` / ``max_range`` / ``initial_threshold`` are the KISS-ICP config values the
        reference reads (config.mapping.voxel_size = max_range / 100, 20, 100 m, config.adaptive_threshold.initial_threshold
        = 2).  ``preprocess=False`` skips the voxel steps: the clouds are matched at the density they are given.

        ``score``: how a RANSAC hypothesis is scored.  "nn_all" = as Open3D 0.18 does it inside the reference's solver call
        (registration_node.py:319-327): nearest map point of EVERY transformed scan point, fitness / rmse over
        those -- with the reference's max_correspondence_distance = 10000 the minimum scan -> map chamfer RMSE wins.
        "corr" = inlier count and residuals over the correspondence list (BASELINE.json's form).  "auto" (default) = "nn_all"
        at the reference's literal distance (>= 10000), "corr" for a real inlier threshold."""
        if score not in ("auto", "nn_all", "corr"):
            raise ValueError(f"Invalid score: {score}")
        self.score = ("nn_all" if max_correspondence_distance >= 1e4 else "corr") if score == "auto" else score
        self.ransac_iters, self.max_dist, self.min_cosine, self.seed, self.device = (ransac_iters, max_correspondence_distance,
                                                                                   min_cosine, seed, device)
        self.voxel_size, self.max_points_per_voxel, self.max_range = voxel_size, max_points_per_voxel, max_range
        self.initial_threshold, self.preprocess = initial_threshold, preprocess
        self.rot_errors, self.trans_errors = {}, {}

    def _voxel_scan(self, raw_scan: np.ndarray) -> np.ndarray:
        # "double-downsampling from KISS-ICP" (registration_node.py:287-288, 399-400)
        return voxel_down_sample(voxel_down_sample(raw_scan, self.voxel_size * 0.5), self.voxel_size * 1.0)

    def _new_map(self) -> VoxelHashMap:
        return VoxelHashMap(self.voxel_size, self.max_range, self.max_points_per_voxel, device=self.device)

    def compute_vfm_correspondences(self, voxel_map: np.ndarray, raw_scan: np.ndarray, initial_pose=np.eye(4)):
        """registration_node.py:396-425: thin the map (<= 20 points per voxel), down-sample the scan (0.5 v, 1 v, then 5 m;
        1 m if that leaves fewer than 75 matches), top-1 cosine >= 0.8."""
        vmap = self._new_map()
        vmap.add_points(voxel_map)
        if not self.preprocess:
            return vmap.get_vfm_correspondences(metrics.transform_pcl(raw_scan, initial_pose), self.min_cosine)
        pcl = metrics.transform_pcl(self._voxel_scan(raw_scan), initial_pose)
        corr = vmap.get_vfm_correspondences(voxel_down_sample(pcl, 5.0), self.min_cosine)
        if corr[0].shape[0] < 75:
            corr = vmap.get_vfm_correspondences(voxel_down_sample(pcl, 1.0), self.min_cosine)
        return corr

    def teaser_registration(self, voxel_map: np.ndarray, raw_scan: np.ndarray, method: str, run_icp: bool = False):
        """registration_node.py:91-160 for method='vfm': descriptor correspondences (compute_vfm_correspondences) -> the
        TEASER++ solve with the reference's solver parameters (noise bound 0.2, PMC_EXACT, CHAIN, GNC-TLS) -> optionally the
        same ICP refinement as ransac_registration.  Returns (teaser_pose, icp_pose | None)."""
        if method != "vfm":
            raise ValueError(f"Invalid method: {method}")   # :108-109 ('fpfh' needs Open3D's FPFH features: not on this path)
        voxel_map, raw_scan = np.asarray(voxel_map), np.asarray(raw_scan)
        if voxel_map.ndim != 2 or raw_scan.ndim != 2 or voxel_map.shape[1] != raw_scan.shape[1] or raw_scan.shape[1] <= 3:
            raise ValueError("Invalid shape")
        src, tgt = self.compute_vfm_correspondences(voxel_map, raw_scan)
        teaser_pose = api.teaser_solve(src, tgt, noise_bound=0.2, cbar2=1.0, gnc_factor=1.4, max_iterations=10000, cost_threshold=1e-16,
                                       device=self.device).T
        if not run_icp:
            return teaser_pose, None
        teaser_pose = teaser_pose.copy()
        teaser_pose[:3, :3] = metrics.orthogonalize_rotation(teaser_pose[:3, :3])   # :143-148
        icp_map = self._new_map()
        icp_map.add_points(np.ascontiguousarray(voxel_map[:, :3]))
        scan_xyz = self._voxel_scan(raw_scan[:, :3]) if self.preprocess else raw_scan[:, :3]
        sigma = self.initial_threshold
        return teaser_pose, register_frame(scan_xyz, icp_map, teaser_pose, 3 * sigma, sigma / 3)   # :150-156

    def ransac_registration(self, voxel_map: np.ndarray, raw_scan: np.ndarray, method: str, run_icp: bool = False):
        """registration_node.py:273-357 for method='vfm': correspondences -> RANSAC (-> ICP refinement).  The reference
        recovers correspondence indices in the voxelised clouds with KD-trees (:289-309); here the voxel operations
        return indices, so the correspondences are first-class and that step disappears."""
        if method != "vfm":
            if method in ("fpfh", "dip", "gedi", "fcgf", "gcl", "spinnet"):
                raise NotImplementedError(f"baseline descriptor '{method}' is outside the VFM hot path")
            raise ValueError(f"Invalid method: {method}")  # registration_node.py:284-285
        voxel_map, raw_scan = np.asarray(voxel_map), np.asarray(raw_scan)
        if voxel_map.ndim != 2 or raw_scan.ndim != 2 or voxel_map.shape[1] != raw_scan.shape[1] or raw_scan.shape[1] <= 3:
            raise ValueError("Invalid shape")
        kw = dict(normalize=True, min_cos=self.min_cosine, ransac_iters=self.ransac_iters, inlier_thresh=self.max_dist,
                  seed=self.seed, device=self.device)
        if self.score == "nn_all":
            ransac_pose, scan_xyz = self._ransac_nn_all(voxel_map, raw_scan)
            r = None
        elif not self.preprocess:
            r = api.register(raw_scan[:, :3], voxel_map[:, :3], raw_scan[:, 3:], voxel_map[:, 3:], **kw)
            scan_xyz, vmap = raw_scan[:, :3], None
        else:
            vmap = self._new_map()
            vmap.add_points(voxel_map)                     # descriptor map, thinned
            res_map = vmap.resident()                      # prepared once for both attempts below
            voxel_scan = self._voxel_scan(raw_scan)        # (n, 3 + D); its xyz is the reference's `voxel_scan`
            scan_xyz = voxel_scan[:, :3]
            r = None
            skw = dict(min_cos=self.min_cosine, ransac_iters=self.ransac_iters, inlier_thresh=self.max_dist, seed=self.seed)
            for leaf in (5.0, 1.0):                        # "Voxelized too sparse, retrying with a larger voxel size"
                q = voxel_down_sample(voxel_scan, leaf)
                r = api.register_scans(res_map, [(np.ascontiguousarray(q[:, :3]), np.ascontiguousarray(q[:, 3:]))], **skw)[0]
                if len(r.corr) >= 75:
                    break
        if r is not None:
            ransac_pose = r.T
        if not run_icp:
            return ransac_pose, None
        ransac_pose = ransac_pose.copy()
        ransac_pose[:3, :3] = metrics.orthogonalize_rotation(ransac_pose[:3, :3])   # registration_node.py:331-336
        icp_map = self._new_map()
        icp_map.add_points(np.ascontiguousarray(voxel_map[:, :3]))                   # :290-291
        sigma = self.initial_threshold
        pose = register_frame(scan_xyz, icp_map, ransac_pose, 3 * sigma, sigma / 3)  # :337-341
        return ransac_pose, pose

    def _ransac_nn_all(self, voxel_map: np.ndarray, raw_scan: np.ndarray):
        """Correspondences as above, solved with the Open3D-style hypothesis score: source = the voxelised scan (`voxel_scan`,
        registration_node.py:287-288), target = the points of the 3-D voxel hash map (`voxel_map_3d`, :290-293), hypotheses
        from the descriptor correspondences (:312-327).  Returns (pose, voxelised scan xyz)."""
        dev = api.get_context(self.device).device
        if not self.preprocess:
            m = api.match_nn(raw_scan[:, 3:], voxel_map[:, 3:], normalize=True, device=dev)
            corr = api.filter_correspondences(m, min_cos=self.min_cosine, device=dev)
            src_xyz, tgt_xyz = raw_scan[:, :3], voxel_map[:, :3]
        else:
            vmap = self._new_map()
            vmap.add_points(voxel_map)
            voxel_scan = self._voxel_scan(raw_scan)
            src_xyz, tgt_xyz = voxel_scan[:, :3], vmap._xyzn          # the same thinning as the 3-D map: same kept points
            corr = None
            for leaf in (5.0, 1.0):                                   # "Voxelized too sparse, retrying ..." (:420-423)
                q, qidx = voxel_down_sample(voxel_scan, leaf, return_index=True)
                m = vmap.resident().match(q[:, 3:], min_cos=self.min_cosine, second=False)
                corr = api.filter_correspondences(m, min_cos=self.min_cosine, device=dev)
                corr[:, 0] = torch.from_numpy(np.asarray(qidx)).to(corr.device, torch.int32)[corr[:, 0].long()]   # rows of voxel_scan
                if corr.shape[0] >= 75:
                    break
        r = api.ransac_nn_all(src_xyz, tgt_xyz, corr, n_hyp=self.ransac_iters, max_dist=self.max_dist, seed=self.seed, device=dev)
        return r.T, src_xyz

    def icp_registration(self, voxel_map: np.ndarray, raw_scan: np.ndarray, initial_pose=None, dist: float = 3):
        """registration_node.py:358-394: double-down-sample the scan, thin the map into a voxel hash map, ICP from
        ``initial_pose`` with max correspondence distance ``dist * sigma`` and kernel ``sigma / dist``.  (N, 3) clouds run the
        point-to-point loop, descriptor-carrying clouds the VFM-ICP overload -- as ``register_frame`` dispatches."""
        voxel_map, raw_scan = np.asarray(voxel_map), np.asarray(raw_scan)
        if voxel_map.ndim != 2 or raw_scan.ndim != 2 or raw_scan.shape[1] < 3 or voxel_map.shape[1] != raw_scan.shape[1]:
            raise ValueError("Invalid shape")
        voxel_scan = self._voxel_scan(raw_scan)
        vmap = self._new_map()
        vmap.add_points(voxel_map)
        sigma = self.initial_threshold
        pose0 = np.eye(4) if initial_pose is None else np.asarray(initial_pose, dtype=np.float64)
        return register_frame(voxel_scan, vmap, pose0, dist * sigma, sigma / dist)

    def compute_errors(self, pose, gt_pose, method: str):
        t, r = metrics.compute_errors(np.asarray(pose), np.asarray(gt_pose))
        self.rot_errors.setdefault(method, []).append(r)
        self.trans_errors.setdefault(method, []).append(t)
        return t, r

    def compute_success_rate(self, method: str, translation_threshold, rotation_threshold) -> float:
        return metrics.success_rate(self.trans_errors[method], self.rot_errors[method], translation_threshold, rotation_threshold)
