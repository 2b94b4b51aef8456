"""Oracle (test infrastructure): pose post-processing and error metrics, float64 NumPy.
PINNED against the reference's own functions (tests/golden/metrics.npz).

Follows (paths relative to /root/reference):
  * ``src/vfm-reg/src/registration_node.py:333-336``   Newton orthogonalisation of R
  * ``src/vfm-reg/src/registration_node.py:997-1019``  compute_errors (RTE, RRE)
  * ``src/vfm-reg/src/registration_node.py:1021-1025`` compute_success_rate
  * ``src/vfm-reg/src/vfm_reg/utils.py:47-54``         transform_pcl
"""
from __future__ import annotations

import numpy as np


def orthogonalize(r: np.ndarray, max_iter: int = 100) -> np.ndarray:
    r = np.array(r, dtype=np.float64)
    it = 0
    while np.abs(1 - np.linalg.det(r)) > 1e-12 and it < max_iter:
        r = 3 / 2 * r - 1 / 2 * r @ r.T @ r
        it += 1
    return r


def compute_errors(pose: np.ndarray, gt_pose: np.ndarray):
    """(trans_error [m], rot_error [deg])."""
    r, r_gt = pose[:3, :3], gt_pose[:3, :3]
    rot = abs(np.arccos(min(max(((r.T @ r_gt).trace() - 1) / 2, -1.0), 1.0)))
    return float(np.linalg.norm(pose[:3, 3] - gt_pose[:3, 3])), float(np.rad2deg(rot))


def success_rate(trans_errors, rot_errors, t_thresh: float, r_thresh: float) -> float:
    t = np.asarray(trans_errors) < t_thresh
    r = np.asarray(rot_errors) < r_thresh
    return float(np.mean(t & r))


def transform_pcl(pcl: np.ndarray, transform: np.ndarray) -> np.ndarray:
    assert transform.shape == (4, 4), "Invalid shape"
    xyz = pcl[:, :3].T.copy()
    xyz = np.insert(xyz, 3, values=1, axis=0)
    xyz = transform @ xyz
    out = np.c_[xyz.T[:, :3], pcl[:, 3:]]
    return out.astype(pcl.dtype)
