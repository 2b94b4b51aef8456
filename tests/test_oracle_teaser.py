"""oracle/teaser.py (the TEASER++-style restatement, parity unpinned) checked through properties: a planted rigid motion is
recovered under heavy outliers, the clique is the planted inlier set, and the scalar TLS estimator ignores outliers."""
import numpy as np

from oracle import teaser as ot


def _problem(seed, n_in, n_out, noise=0.02):
    rng = np.random.default_rng(seed)
    from scipy.spatial.transform import Rotation as R
    rot = R.from_euler("zyx", rng.uniform(-180, 180, 3) * [1, 0.05, 0.05], degrees=True).as_matrix()
    t = rng.normal(0, 10, 3)
    src = rng.uniform(-40, 40, (n_in + n_out, 3))
    tgt = src @ rot.T + t + rng.normal(0, noise, (n_in + n_out, 3))
    tgt[n_in:] = rng.uniform(-40, 40, (n_out, 3))
    perm = rng.permutation(n_in + n_out)
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = rot, t
    return src[perm], tgt[perm], T, np.sort(np.argsort(perm)[:n_in])


def test_planted_motion_is_recovered_under_outliers():
    src, tgt, T, inl = _problem(1, 40, 110)
    est, cl = ot.teaser_solve(src, tgt)
    assert set(inl.tolist()) <= set(cl.tolist()) and len(cl) <= len(inl) + 2      # the inliers form the maximum clique
    assert np.linalg.norm(est[:3, 3] - T[:3, 3]) < 0.05
    assert np.degrees(np.arccos(np.clip((np.trace(est[:3, :3].T @ T[:3, :3]) - 1) / 2, -1, 1))) < 0.2


def test_tls_scalar_ignores_outliers():
    rng = np.random.default_rng(3)
    x = np.r_[5.0 + rng.normal(0, 0.05, 30), rng.uniform(-50, 50, 60)]
    assert abs(ot.tls_scalar(x, 0.2) - 5.0) < 0.05


def test_degenerate_inputs():
    T, cl = ot.teaser_solve(np.zeros((1, 3)), np.zeros((1, 3)))
    assert np.array_equal(T, np.eye(4)) and len(cl) == 0
