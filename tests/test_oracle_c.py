"""The canonical-order C restatement (oracle/c) against the float64 NumPy oracle, and known-answer cases
(SURVEY.md section 8c): identity, pure translation, planted SE(3) with outliers, reflection-forcing and
degenerate (collinear / repeated) triplets, zero-norm descriptors, K < 3."""
import numpy as np
import pytest

from oracle import cref, match, ransac
from vfm_registration_b200 import synth


def test_renorm_and_top2_agree_with_numpy():
    rng = np.random.default_rng(1)
    a = rng.standard_normal((300, 96)).astype(np.float32)
    b = rng.standard_normal((700, 96)).astype(np.float32)
    a[5] = 0  # zero-norm query row stays zero (faiss `if (nr > 0)`)
    b[7] = 0
    an = cref.renorm_l2(a)
    assert np.all(an[5] == 0)
    assert np.abs(np.linalg.norm(np.delete(an, 5, 0).astype(np.float64), axis=1) - 1).max() < 1e-6
    assert np.abs(an - match.renorm_l2(a)).max() < 1e-6
    rc = cref.match_nn(a, b, mutual=True)
    ro = match.match_nn(a, b, mutual=True)
    for k in ("01", "10"):
        gap = ro["sim" + k] - ro["sec" + k]
        clear = gap > 1e-5
        assert clear.mean() > 0.95
        assert np.array_equal(rc["idx" + k][clear], ro["idx" + k][clear])
        assert np.abs(rc["sim" + k] - ro["sim" + k]).max() < 1e-5
        assert np.abs(rc["sec" + k] - ro["sec" + k]).max() < 1e-5
    # zero query: every inner product is 0 -> ties -> lowest index, similarity 0 (rejected by the 0.8 gate)
    assert rc["idx01"][5] == 0 and rc["sim01"][5] == 0 and ro["idx01"][5] == 0


def test_exact_ties_pick_lowest_index():
    rng = np.random.default_rng(2)
    b = rng.standard_normal((64, 32)).astype(np.float32)
    b[40] = b[3]
    b[20] = b[3]
    a = b[[3, 20, 40, 10]].copy()
    r = cref.match_nn(a, b)
    assert list(r["idx01"]) == [3, 3, 3, 10]
    assert np.all(r["sec01"][:3] == r["sim01"][:3])  # runner-up of a tie is the tied value


def test_sampler_matches_numpy():
    for seed, h, k in ((0, 100, 7), (42, 5000, 8191), (2 ** 63 + 5, 64, 3)):
        assert np.array_equal(cref.sample_indices(seed, h, k), ransac.sample_indices(seed, h, k))
        assert cref.sample_indices(seed, h, k).max() < k


def _rigid(rng):
    from scipy.spatial.transform import Rotation as R
    return R.from_rotvec(rng.normal(0, 1.0, 3)).as_matrix(), rng.normal(0, 5, 3)


def test_kabsch3_known_answers():
    rng = np.random.default_rng(3)
    p = rng.uniform(-10, 10, (3, 3))
    r, t, ok = cref.kabsch3(p, p)
    assert ok and np.abs(r - np.eye(3)).max() < 1e-12 and np.abs(t).max() < 1e-12
    r, t, ok = cref.kabsch3(p, p + [1.0, -2.0, 0.5])
    assert ok and np.abs(r - np.eye(3)).max() < 1e-12 and np.abs(t - [1.0, -2.0, 0.5]).max() < 1e-12
    for _ in range(50):
        rg, tg = _rigid(rng)
        p = rng.uniform(-30, 30, (3, 3))
        r, t, ok = cref.kabsch3(p, p @ rg.T + tg)
        assert ok and np.abs(r - rg).max() < 1e-9 and np.abs(t - tg).max() < 1e-8
        ro, to, oko = ransac.kabsch(p, p @ rg.T + tg)
        assert oko and np.abs(r - ro).max() < 1e-9
    # 90 degree rotation about z
    rz = np.array([[0.0, -1, 0], [1, 0, 0], [0, 0, 1]])
    p = np.array([[1.0, 0, 0], [0, 2.0, 0], [0, 0, 3.0]])
    r, t, ok = cref.kabsch3(p, p @ rz.T)
    assert ok and np.abs(r - rz).max() < 1e-12
    # reflection-forcing target (mirror): result must still be a proper rotation, equal to the SVD rule
    q = (p @ rz.T) * np.array([1.0, 1.0, -1.0]) + rng.normal(0, 0.1, (3, 3))
    r, t, ok = cref.kabsch3(p, q)
    ro, to, _ = ransac.kabsch(p, q)
    assert ok and abs(np.linalg.det(r) - 1) < 1e-12 and np.abs(r - ro).max() < 1e-9 and np.abs(t - to).max() < 1e-9


def test_kabsch3_degenerate():
    p = np.array([[0.0, 0, 0], [1, 1, 1], [2, 2, 2]])          # collinear
    assert not cref.kabsch3(p, p + 1.0)[2]
    p = np.array([[1.0, 2, 3], [1, 2, 3], [4, 5, 6]])          # repeated sample (drawn with replacement)
    assert not cref.kabsch3(p, p)[2]
    p = np.array([[1.0, 2, 3], [1, 2, 3], [1, 2, 3]])          # all equal
    r, t, ok = cref.kabsch3(p, p)
    assert not ok and np.all(np.isfinite(r))
    assert not ransac.kabsch(p, p)[2]


@pytest.mark.parametrize("thresh", [1.0, 1e4])
def test_ransac_c_vs_numpy(thresh):
    s = synth.make_pair(11, 3000, 1200, 32, inlier_frac=0.3)
    k = 600
    rng = np.random.default_rng(5)
    good = np.nonzero(s["perm"] >= 0)[0]
    bad = np.nonzero(s["perm"] < 0)[0]
    q = np.concatenate([good[:180], bad[: k - 180]])
    q.sort()
    corr = np.stack([q, np.where(s["perm"][q] >= 0, s["perm"][q], rng.integers(0, 3000, len(q)))], 1).astype(np.int32)
    si = ransac.sample_indices(7, 4096, k)
    rc = cref.ransac(s["scan_xyz"], s["map_xyz"], corr, si, thresh)
    ro = ransac.ransac(s["scan_xyz"], s["map_xyz"], corr, si, thresh)
    assert np.array_equal(rc["counts"], ro["counts"])
    assert rc["best"] == ro["best"] and np.array_equal(rc["mask"], ro["mask"])
    assert np.linalg.norm(rc["T"] - ro["T"]) < 1e-9
    if thresh == 1.0:
        rte, rre = synth.pose_errors(rc["T"], s["T_gt"])
        assert rte < 0.3 and rre < 1.0 and rc["n_inliers"] >= 150
    else:  # the reference's literal tau = 10000: everything is an inlier, winner = min residual sum
        assert rc["n_inliers"] == k and rc["fitness"] == 1.0
        assert rc["best"] == int(np.argmin(np.where(ro["valid"], ro["sumq"], np.iinfo(np.int64).max)))
    # device-RNG path == injected indices of the same seed
    r2 = cref.ransac(s["scan_xyz"], s["map_xyz"], corr, None, thresh, seed=7, n_hyp=4096)
    assert r2["best"] == rc["best"] and np.array_equal(r2["T"], rc["T"])
    # refit
    r3 = cref.ransac(s["scan_xyz"], s["map_xyz"], corr, si, thresh, refit=True)
    o3 = ransac.ransac(s["scan_xyz"], s["map_xyz"], corr, si, thresh, refit=True)
    assert np.linalg.norm(r3["T"] - o3["T"]) < 1e-9


def test_ransac_small_k():
    src = np.zeros((5, 3))
    for k in (0, 1, 2):
        corr = np.stack([np.arange(k), np.arange(k)], 1).astype(np.int32)
        si = np.zeros((16, 3), dtype=np.int32)
        for impl in (cref.ransac, ransac.ransac):
            r = impl(src, src, corr, si, 1.0)
            assert r["best"] == -1 and np.array_equal(r["T"], np.eye(4)) and r["fitness"] == 0.0


def test_get_vfm_correspondences_layout():
    s = synth.make_pair(3, 800, 300, 48)
    pts = np.c_[s["scan_xyz"], s["scan_feat"]].astype(np.float64)
    mp = np.c_[s["map_xyz"], s["map_feat"]].astype(np.float64)
    src, tgt = match.get_vfm_correspondences(pts, mp, 0.8)
    inl = np.nonzero(s["perm"] >= 0)[0]
    assert src.dtype == np.float64 and src.shape == tgt.shape
    assert len(src) == len(inl)  # planted matches have cos ~0.9, random pairs ~0
    assert np.array_equal(src, s["scan_xyz"][inl].astype(np.float64))
    assert np.array_equal(tgt, s["map_xyz"][s["perm"][inl]].astype(np.float64))
    # nothing passes a gate of 1.1 -> K = 0, no crash (the reference would hit UB, VoxelHashMap.cpp:547-550)
    e0, e1 = match.get_vfm_correspondences(pts, mp, 1.1)
    assert e0.shape == (0, 3) and e1.shape == (0, 3)


def test_nn_all_oracle_known_answer():
    """The Open3D-style score (oracle/ransac.py:ransac_nn_all): with max_dist = 1e4 every point counts, the planted pose wins,
    and the score of a hypothesis equals a brute-force chamfer evaluation."""
    from oracle import ransac
    from vfm_registration_b200 import synth
    s = synth.make_pair(5, 1500, 400, 8, inlier_frac=0.6)
    good = np.nonzero(s["perm"] >= 0)[0][:150]
    corr = np.stack([good, s["perm"][good]], 1).astype(np.int32)
    si = ransac.sample_indices(1, 64, len(corr))
    o = ransac.ransac_nn_all(s["scan_xyz"], s["map_xyz"], corr, si, 1e4)
    assert o["fitness"] == 1.0 and (o["inliers"][o["inliers"] >= 0] == 400).all()
    rte, rre = synth.pose_errors(o["T"], s["T_gt"])
    assert rte < 0.5 and rre < 2.0
    h = o["best"]
    x = s["scan_xyz"].astype(np.float64) @ o["T"][:3, :3].T + o["T"][:3, 3]
    d2 = ((x[:, None, :] - s["map_xyz"].astype(np.float64)[None]) ** 2).sum(-1).min(1)
    assert abs(d2.sum() - o["sum_d2"][h]) < 1e-9 * max(1.0, d2.sum())
    assert o["sum_d2"][h] == min(v for v, c in zip(o["sum_d2"], o["inliers"]) if c > 0)
