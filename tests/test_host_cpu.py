"""Host-side logic that needs no GPU: pose utilities against the reference-generated golden fixture, synthetic data."""
import numpy as np

from vfm_registration_b200 import metrics, synth


def test_metrics_match_reference_golden(golden):
    g = golden("metrics.npz")
    errs = np.array([metrics.compute_errors(p, q) for p, q in zip(g["poses"], g["gts"])])
    assert np.array_equal(errs, g["errs"])
    rates = [metrics.success_rate(errs[:, 0], errs[:, 1], t, r) for t, r in ((0.3, 15), (0.6, 1.5), (2, 5), (1, 5))]
    assert np.array_equal(np.array(rates), g["rates"])
    got = metrics.transform_pcl(g["pcl"], g["poses"][3])
    assert got.dtype == g["pcl_t"].dtype and np.abs(got - g["pcl_t"]).max() < 1e-5  # f32 output, matmul order differs


def test_orthogonalize():
    rng = np.random.default_rng(0)
    from scipy.spatial.transform import Rotation as R
    r = R.from_rotvec(rng.normal(0, 1, 3)).as_matrix() + rng.normal(0, 1e-4, (3, 3))
    o = metrics.orthogonalize_rotation(r)
    assert abs(np.linalg.det(o) - 1) <= 1e-12 and np.abs(o @ o.T - np.eye(3)).max() < 1e-9


def test_synth_pair_is_consistent():
    s = synth.make_pair(5, 2000, 800, 384)
    inl = np.nonzero(s["perm"] >= 0)[0]
    assert len(inl) == 240
    back = s["scan_xyz"][inl].astype(np.float64) @ s["T_gt"][:3, :3].T + s["T_gt"][:3, 3]
    assert np.abs(back - s["map_xyz"][s["perm"][inl]]).max() < 0.15
    cos = (s["scan_feat"][inl] * s["map_feat"][s["perm"][inl]]).sum(1)
    assert 0.85 < cos.mean() < 0.95 and cos.min() > 0.8
    assert synth.pose_errors(s["T_gt"], s["T_gt"]) == (0.0, 0.0)
