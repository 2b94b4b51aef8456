"""Synthetic scene pairs of NCLT / RobotCar shape: there are no datasets in the container,
so tests and bench.py draw (map, scan) pairs with a planted SE(3), a known inlier fraction and random unit
descriptors whose true matches have cosine ~0.9 (above the reference's 0.8 gate, registration_node.py:418)."""
from __future__ import annotations

import numpy as np


def random_pose(rng: np.random.Generator) -> np.ndarray:
    """Ground-truth pose like the reference's noise model (registration_node.py:847-853): free yaw, small roll/pitch,
    N(0, 10 m) planar and N(0, 1 m) vertical translation."""
    yaw = rng.uniform(-np.pi, np.pi)
    roll, pitch = np.deg2rad(rng.normal(0.0, 2.0, 2))
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]])
    ry = np.array([[cp, 0, sp], [0, 1.0, 0], [-sp, 0, cp]])
    rx = np.array([[1.0, 0, 0], [0, cr, -sr], [0, sr, cr]])
    t = np.eye(4)
    t[:3, :3] = rz @ ry @ rx
    t[:3, 3] = [rng.normal(0, 10.0), rng.normal(0, 10.0), rng.normal(0, 1.0)]
    return t


def _unit_rows(rng, n, d):
    f = rng.standard_normal((n, d)).astype(np.float32)
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    return f


_LO, _HI = np.array([-50.0, -50.0, -2.0]), np.array([50.0, 50.0, 8.0])


def _make_scan(rng, map_xyz, map_feat, n_scan, inlier_frac, sigma_f, noise):
    """One scan of a map: planted pose, an inlier fraction copied from the map (+ noise), uniform outliers."""
    n_map, d = map_feat.shape if map_feat is not None else (map_xyz.shape[0], 0)
    t_gt = random_pose(rng)
    t_inv = np.linalg.inv(t_gt)
    n_in = min(int(round(inlier_frac * n_scan)), n_map)
    perm = np.full(n_scan, -1, dtype=np.int64)
    which = rng.permutation(n_scan)[:n_in]
    perm[which] = rng.permutation(n_map)[:n_in]
    scan_xyz = rng.uniform(_LO, _HI, (n_scan, 3))
    src = map_xyz[perm[which]]
    scan_xyz[which] = src @ t_inv[:3, :3].T + t_inv[:3, 3] + rng.normal(0.0, noise, (n_in, 3))
    return t_gt, perm, which, n_in, scan_xyz


def make_pair(seed: int, n_map: int, n_scan: int, d: int, inlier_frac: float = 0.3, sigma_f: float = 0.025,
              noise: float = 0.02, scan_seed=None) -> dict:
    """Returns float32 arrays map_xyz (M,3), scan_xyz (N,3), map_feat (M,D), scan_feat (N,D) and the float64 pose
    ``T_gt`` that maps the scan into the map; ``perm[i]`` is the map index scan point i was copied from (-1 = outlier).

    ``scan_seed``: draw the scan (pose, inlier choice, outliers, descriptor noise) from another stream while the map stays
    the one of ``seed`` -- the scans of one scene share their map (registration_node.py:554-590); see ``make_scene``."""
    rng = np.random.default_rng(seed)
    map_xyz = rng.uniform(_LO, _HI, (n_map, 3))
    map_feat = None
    if scan_seed is not None:
        map_feat = _unit_rows(np.random.default_rng([seed, 1]), n_map, d)
        rng = np.random.default_rng([seed, 2, scan_seed])
    t_gt, perm, which, n_in, scan_xyz = _make_scan(rng, map_xyz, map_feat, n_scan, inlier_frac, sigma_f, noise)
    if map_feat is None:
        map_feat = _unit_rows(rng, n_map, d)
    scan_feat = rng.standard_normal((n_scan, d)).astype(np.float32)
    scan_feat[which] = map_feat[perm[which]] + sigma_f * rng.standard_normal((n_in, d)).astype(np.float32)
    scan_feat /= np.linalg.norm(scan_feat, axis=1, keepdims=True)
    return dict(map_xyz=map_xyz.astype(np.float32), scan_xyz=scan_xyz.astype(np.float32), map_feat=map_feat,
                scan_feat=scan_feat.astype(np.float32), T_gt=t_gt, perm=perm)


def make_scene(seed: int, n_map: int, n_scans: int, n_scan: int, d: int, inlier_frac: float = 0.3, sigma_f: float = 0.025,
               noise: float = 0.02) -> dict:
    """One map and ``n_scans`` scans of it, as the reference's scenes are built (one local map, 5 scans per NCLT scene and 3
    per RobotCar scene, data/*/scene_*.json): ``scans[k]`` equals ``make_pair(seed, ..., scan_seed=k)`` without its map."""
    map_xyz = np.random.default_rng(seed).uniform(_LO, _HI, (n_map, 3))
    map_feat = _unit_rows(np.random.default_rng([seed, 1]), n_map, d)
    scans = []
    for k in range(n_scans):
        rng = np.random.default_rng([seed, 2, k])
        t_gt, perm, which, n_in, scan_xyz = _make_scan(rng, map_xyz, map_feat, n_scan, inlier_frac, sigma_f, noise)
        scan_feat = rng.standard_normal((n_scan, d)).astype(np.float32)
        scan_feat[which] = map_feat[perm[which]] + sigma_f * rng.standard_normal((n_in, d)).astype(np.float32)
        scan_feat /= np.linalg.norm(scan_feat, axis=1, keepdims=True)
        scans.append(dict(scan_xyz=scan_xyz.astype(np.float32), scan_feat=scan_feat.astype(np.float32), T_gt=t_gt, perm=perm))
    return dict(map_xyz=map_xyz.astype(np.float32), map_feat=map_feat, scans=scans)


def pose_errors(t: np.ndarray, t_gt: np.ndarray):
    """(RTE [m], RRE [deg]) -- the reference's formulas, registration_node.py:997-1019."""
    r = t[:3, :3].T @ t_gt[:3, :3]
    rre = np.rad2deg(abs(np.arccos(min(max((np.trace(r) - 1) / 2, -1.0), 1.0))))
    return float(np.linalg.norm(t[:3, 3] - t_gt[:3, 3])), float(rre)
