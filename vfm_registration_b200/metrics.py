"""Host-side pose utilities of the path's tail: negligible float64 arithmetic that the reference
also does in NumPy.  Reference: registration_node.py:333-336 (Newton orthogonalisation), :997-1025 (RTE / RRE /
success rate), vfm_reg/utils.py:47-54 (transform_pcl)."""
from __future__ import annotations

import numpy as np


def orthogonalize_rotation(r: np.ndarray, max_iter: int = 100) -> np.ndarray:
    """R <- 3/2 R - 1/2 R R^T R until |1 - det R| <= 1e-12 (registration_node.py:333-336)."""
    r = np.array(r, dtype=np.float64)
    for _ in range(max_iter):
        if np.abs(1 - np.linalg.det(r)) <= 1e-12:
            break
        r = 3 / 2 * r - 1 / 2 * r @ r.T @ r
    return r


def compute_errors(pose: np.ndarray, gt_pose: np.ndarray):
    """(translation error [m], rotation error [deg]) exactly as registration_node.py:997-1019."""
    r, r_gt = pose[:3, :3], gt_pose[:3, :3]
    rot = abs(np.arccos(min(max(((r.T @ r_gt).trace() - 1) / 2, -1.0), 1.0)))
    return float(np.linalg.norm(pose[:3, 3] - gt_pose[:3, 3])), float(np.rad2deg(rot))


def success_rate(trans_errors, rot_errors, translation_threshold: float, rotation_threshold: float) -> float:
    """registration_node.py:1021-1025.  The reference reports (0.3 m, 15 deg), (0.6 m, 1.5 deg), (2 m, 5 deg);
    BASELINE.json adds (1 m, 5 deg)."""
    ok = (np.asarray(trans_errors) < translation_threshold) & (np.asarray(rot_errors) < rotation_threshold)
    return float(np.mean(ok))


def transform_pcl(pcl: np.ndarray, transform: np.ndarray) -> np.ndarray:
    """T [xyz; 1] keeping the extra (descriptor) columns, cast back to the input dtype (vfm_reg/utils.py:47-54)."""
    assert transform.shape == (4, 4), "Invalid shape"
    xyz = pcl[:, :3].astype(np.float64) @ transform[:3, :3].T + transform[:3, 3]
    out = np.c_[xyz, pcl[:, 3:]]
    assert out.shape == pcl.shape
    return out.astype(pcl.dtype)
