"""CPU checks of the voxel / ICP oracle (oracle/voxel.py).  The reference's C++ for these rows cannot be built offline and
has no golden vectors (parity unpinned, see the oracle header), so the restatement is checked through properties:
brute-force nearest neighbours, scipy's matrix exponential, recovery of a planted SE(3)."""
import numpy as np
import pytest
from scipy.linalg import expm
from scipy.spatial import cKDTree

from oracle import voxel as ov


def _cloud(seed, n, span=20.0):
    rng = np.random.default_rng(seed)
    return rng.uniform(-span, span, (n, 3)) * np.array([1.0, 1.0, 0.2])


def test_voxel_index_truncates_toward_zero():
    p = np.array([[-0.4, 0.4, -1.0], [-1.2, 1.99, 2.0], [-0.0, 3.5, -3.5]])
    assert ov.voxel_index(p, 1.0).tolist() == [[0, 0, -1], [-1, 1, 2], [0, 3, -3]]
    assert ov.voxel_index(p, 0.5).tolist() == [[0, 0, -2], [-2, 3, 4], [0, 7, -7]]


def test_voxel_down_sample_first_in_voxel():
    pts = _cloud(1, 5000)
    out, idx = ov.voxel_down_sample(pts, 1.0, return_index=True)
    vox = ov.voxel_index(pts, 1.0)
    _, first = np.unique(vox, axis=0, return_index=True)
    assert np.array_equal(np.sort(first), idx) and np.array_equal(out, pts[idx])
    # the (-1, 1) double-width voxel around zero
    q = np.array([[-0.9, 0.1, 0.1], [0.9, 0.1, 0.1], [1.1, 0.1, 0.1]])
    assert len(ov.voxel_down_sample(q, 1.0)) == 2


def test_map_keeps_first_points_per_voxel_across_calls():
    pts = _cloud(2, 4000, span=4.0)
    m = ov.VoxelHashMapOracle(1.0, 5)
    m.add_points(pts[:2500])
    m.add_points(pts[2500:])
    xyz, ids = m.point_cloud()
    vox = ov.voxel_index(pts, 1.0)
    want = []
    counts = {}
    for i, v in enumerate(map(tuple, vox)):
        if counts.get(v, 0) < 5:
            counts[v] = counts.get(v, 0) + 1
            want.append(i)
    assert np.array_equal(ids, np.asarray(want)) and np.array_equal(xyz, pts[want])


def test_closest_neighbor_matches_brute_force_when_close():
    pts = _cloud(3, 3000)
    m = ov.VoxelHashMapOracle(1.0, 1000)
    m.add_points(pts)
    tree = cKDTree(pts)
    rng = np.random.default_rng(0)
    for q in pts[rng.integers(0, len(pts), 200)] + rng.normal(0, 0.2, (200, 3)):
        p, d2 = m.closest_neighbor(q)
        d, j = tree.query(q)
        if d < 1.0:   # the true neighbour is then inside the 27-voxel neighbourhood
            assert p is not None and np.array_equal(p, pts[j]) and abs(np.sqrt(d2) - d) < 1e-12


def test_se3_exp_matches_matrix_exponential():
    rng = np.random.default_rng(4)
    for scale in (1e-8, 1e-3, 0.3, 2.0):
        dx = rng.normal(0, scale, 6)
        G = np.zeros((4, 4))
        G[:3, :3] = ov.hat(dx[3:])
        G[:3, 3] = dx[:3]
        assert np.abs(ov.se3_exp(dx) - expm(G)).max() < 1e-12


@pytest.mark.parametrize("seed", [5, 6])
def test_register_frame_recovers_planted_pose(seed):
    rng = np.random.default_rng(seed)
    map_pts = _cloud(seed, 4000)
    ang = np.deg2rad(3.0)
    T = np.eye(4)
    T[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]
    T[:3, 3] = [0.3, -0.2, 0.05]
    scan = (map_pts[rng.choice(len(map_pts), 1500, replace=False)] - T[:3, 3]) @ T[:3, :3]   # T^-1 applied
    m = ov.VoxelHashMapOracle(1.0, 20)
    m.add_points(map_pts)
    est, info = ov.register_frame(scan, m, np.eye(4), 3.0, 2.0 / 3.0, return_info=True)
    assert info["iterations"] < 60 and np.abs(est - T).max() < 1e-6
    assert np.array_equal(ov.register_frame(scan, ov.VoxelHashMapOracle(1.0), T, 3.0, 0.5), T)   # empty map -> initial guess


def test_vfm_icp_first_loop_prunes_outliers_and_converges():
    rng = np.random.default_rng(9)
    tgt = _cloud(9, 400)
    ang = np.deg2rad(5.0)
    T = np.eye(4)
    T[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]
    T[:3, 3] = [1.0, -0.5, 0.1]
    src = (tgt - T[:3, 3]) @ T[:3, :3]
    src[:40] += rng.normal(0, 5.0, (40, 3))          # wrong descriptor matches
    assert ov._median_nth([3.0, 1.0, 2.0]) == 2.0 and ov._median_nth([4.0, 1.0, 3.0, 2.0]) == 2.5
    est, j, (s, t) = ov.vfm_icp_first_loop(src, tgt, np.eye(4), 2.0 / 3.0)
    assert 1 <= j < 50 and len(s) <= 365 and np.abs(est - T).max() < 0.05
    e0, j0, (s0, _) = ov.vfm_icp_first_loop(np.zeros((0, 3)), np.zeros((0, 3)), T, 0.5)
    assert j0 == 0 and len(s0) == 0 and np.array_equal(e0, T)


def test_vectorised_restatements_match_the_loops():
    pts = np.concatenate([_cloud(21, 20000, span=6.0), _cloud(22, 50, span=0.4)], axis=0).astype(np.float32)
    for vs in (0.25, 1.0):
        assert np.array_equal(ov.voxel_down_sample_index_fast(pts, vs), ov.voxel_down_sample(pts, vs, return_index=True)[1])
    for cap in (1, 5, 20):
        m = ov.VoxelHashMapOracle(1.0, cap)
        m.add_points(pts)
        assert np.array_equal(ov.voxel_map_kept_index_fast(pts, 1.0, cap), m.point_cloud()[1])
