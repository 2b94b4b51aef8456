"""GPU parity of the voxel operations and the ICP refinement against oracle/voxel.py, called
through the C ABI.  Index sets and nearest neighbours are bit-exact; the ICP pose is float64 Gauss-Newton with a different
(fixed) summation order than the oracle's, tolerance 1e-9 on the 4x4 written in the test."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import voxel as ov  # noqa: E402


@pytest.fixture(scope="module")
def vfm():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import vfm_registration_b200 as v
    v.get_context(0)
    return v


def _cloud(seed, n, span=30.0, dtype=np.float64):
    rng = np.random.default_rng(seed)
    return (rng.uniform(-span, span, (n, 3)) * np.array([1.0, 1.0, 0.15])).astype(dtype)


@pytest.mark.parametrize("n,cols,dtype,vs", [(20000, 3, np.float64, 1.0), (50000, 3, np.float32, 0.5), (3000, 387, np.float32, 1.0),
                                             (1, 3, np.float64, 1.0), (5000, 4, np.float64, 5.0), (1500, 35, np.float64, 0.25)])
def test_voxel_down_sample_vs_oracle(vfm, n, cols, dtype, vs):
    rng = np.random.default_rng(n + cols)
    pts = np.concatenate([_cloud(n, n, dtype=dtype), rng.standard_normal((n, cols - 3)).astype(dtype)], axis=1)
    if n > 10:
        pts[7, :3] = pts[2, :3]                     # duplicate point: the lower index wins
        pts[11, :3] = [-0.3 * vs, 0.2 * vs, 0.0]    # both sides of zero share voxel 0 (truncation toward zero)
        pts[12, :3] = [0.3 * vs, -0.2 * vs, 0.0]
    got, gi = vfm.voxel_down_sample(pts, vs, return_index=True)
    want, wi = ov.voxel_down_sample(pts, vs, return_index=True)
    assert np.array_equal(gi, wi) and got.dtype == pts.dtype and np.array_equal(got, want)
    if n > 10:
        assert 12 not in gi and 7 not in gi
    dev = vfm.voxel_down_sample(torch.from_numpy(pts).cuda(), vs)
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), want)


def test_voxel_down_sample_errors(vfm):
    with pytest.raises(ValueError, match="Invalid shape"):
        vfm.voxel_down_sample(np.zeros((10, 2)), 1.0)
    bad = np.zeros((4, 3))
    bad[2, 0] = np.inf
    with pytest.warns(RuntimeWarning, match="1 of 4"):   # dropped by the host wrapper (the library alone rejects the cloud)
        out, idx = vfm.voxel_down_sample(bad, 1.0, return_index=True)
    assert out.shape == (1, 3) and idx.tolist() == [0]
    with pytest.warns(RuntimeWarning, match="4 of 4"):
        assert vfm.voxel_down_sample(np.full((4, 3), 3e7), 1.0).shape == (0, 3)   # |index| >= 2^20
    assert vfm.voxel_down_sample(np.zeros((0, 3)), 1.0).shape == (0, 3)


@pytest.mark.parametrize("n,vs,cap", [(30000, 1.0, 20), (8000, 2.0, 3), (500, 0.1, 20), (4000, 1.0, 1)])
def test_voxel_map_build_and_nearest_vs_oracle(vfm, n, vs, cap):
    pts = _cloud(10 + cap, n, span=12.0)
    m = vfm.VoxelMap(vs, cap)
    m.build(pts)
    o = ov.VoxelHashMapOracle(vs, cap)
    o.add_points(pts)
    oxyz, oid = o.point_cloud()
    xyz, idx = m.points()
    assert len(m) == len(oid) and np.array_equal(idx.cpu().numpy(), oid) and np.array_equal(xyz.cpu().numpy(), oxyz)
    rng = np.random.default_rng(1)
    q = pts[rng.integers(0, n, 400)] + rng.normal(0, 0.3 * vs, (400, 3))
    q[0] = [1e3, 1e3, 1e3]          # nothing around it
    q[1] = pts[5]                   # exact hit
    tgt, valid, d2 = m.nearest(q, 0.8 * vs)
    tgt, valid, d2 = tgt.cpu().numpy(), valid.cpu().numpy(), d2.cpu().numpy()
    for k in range(len(q)):
        p, od2 = o.closest_neighbor(q[k])
        if p is None:
            assert not valid[k] and d2[k] == -1.0
            continue
        assert d2[k] == od2 and valid[k] == (np.sqrt(od2) < 0.8 * vs)
        if valid[k]:
            assert np.array_equal(tgt[k], p)
    m.close()


@pytest.mark.parametrize("seed,n_scan,noise", [(1, 3000, 0.0), (2, 6000, 0.02), (3, 800, 0.05)])
def test_register_frame_vs_oracle(vfm, seed, n_scan, noise):
    rng = np.random.default_rng(seed)
    map_pts = _cloud(seed, 12000, span=25.0)
    ang = np.deg2rad(rng.uniform(-4, 4))
    T = np.eye(4)
    T[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]
    T[:3, 3] = rng.normal(0, 0.3, 3) * [1, 1, 0.2]
    scan = (map_pts[rng.choice(len(map_pts), n_scan, replace=False)] - T[:3, 3]) @ T[:3, :3] + rng.normal(0, noise, (n_scan, 3))
    guess = np.eye(4)
    guess[:3, 3] = [0.05, -0.05, 0.0]
    m = vfm.VoxelMap(1.0, 20)
    m.build(map_pts)
    o = ov.VoxelHashMapOracle(1.0, 20)
    o.add_points(map_pts)
    got, gi = vfm.register_frame(scan, m, guess, 3.0, 2.0 / 3.0, return_info=True)
    want, wi = ov.register_frame(scan, o, guess, 3.0, 2.0 / 3.0, return_info=True)
    assert gi["iterations"] == wi["iterations"] and gi["correspondences"] == wi["correspondences"]
    assert np.abs(got - want).max() < 1e-9          # float64, different summation order
    assert np.abs(got - T).max() < (1e-6 if noise == 0 else 0.02)
    one, info = vfm.register_frame(scan, m, guess, 3.0, 2.0 / 3.0, max_iterations=1, return_info=True)
    assert info["iterations"] == 1 and np.abs(one - ov.register_frame(scan, o, guess, 3.0, 2.0 / 3.0, max_iterations=1)).max() < 1e-12
    # empty map -> the initial guess; no correspondence -> loop exits with the pose so far
    e = vfm.VoxelMap(1.0, 20)
    assert np.array_equal(vfm.register_frame(scan, e, guess, 3.0, 0.5), guess)
    far = scan + 1e4
    assert np.array_equal(vfm.register_frame(far, m, np.eye(4), 3.0, 0.5), np.eye(4))
    with pytest.raises(ValueError, match="Invalid shape"):
        vfm.register_frame(np.zeros((5, 4)), m, guess, 3.0, 0.5)


def test_compat_ransac_registration_with_icp_vs_oracle_chain(vfm):
    """The reference's ransac_registration(method='vfm', run_icp=True) call surface (registration_node.py:273-357) through
    the shim, against the same chain assembled from the oracles: voxel thinning / down-sampling, top-1 cosine gate, RANSAC,
    Newton orthogonalisation, ICP."""
    from oracle import cref, match
    from vfm_registration_b200 import compat, metrics, synth
    s = synth.make_pair(21, 6000, 3000, 64)
    vmap_arr = np.c_[s["map_xyz"], s["map_feat"]].astype(np.float32)
    scan_arr = np.c_[s["scan_xyz"], s["scan_feat"]].astype(np.float32)
    vmap_arr = np.concatenate([vmap_arr, np.repeat(vmap_arr[:3], 30, axis=0)], axis=0)   # > 20 points in three voxels
    node = compat.RegistrationNode(ransac_iters=4096, max_correspondence_distance=1.0, seed=5)
    ransac_pose, icp_pose = node.ransac_registration(vmap_arr, scan_arr, "vfm", run_icp=True)
    # oracle chain
    om = ov.VoxelHashMapOracle(1.0, 20)
    om.add_points(vmap_arr[:, :3])
    kept_xyz, kept_id = om.point_cloud()
    assert len(kept_id) == 6000 + 3 * 19
    voxel_scan = ov.voxel_down_sample(ov.voxel_down_sample(scan_arr, 0.5), 1.0)
    q = ov.voxel_down_sample(voxel_scan, 5.0)
    mm = cref.match_nn(q[:, 3:], vmap_arr[kept_id, 3:])
    corr = match.filter_correspondences(mm["idx01"], mm["sim01"], min_cos=0.8)
    assert len(corr) >= 75
    c = cref.ransac(np.ascontiguousarray(q[:, :3]), kept_xyz.astype(np.float32), corr, None, 1.0, seed=5, n_hyp=4096)
    want_ransac = c["T"].copy()
    want_ransac[:3, :3] = metrics.orthogonalize_rotation(want_ransac[:3, :3])
    assert np.array_equal(ransac_pose, want_ransac)
    want_icp = ov.register_frame(voxel_scan[:, :3].astype(np.float64), om, want_ransac, 6.0, 2.0 / 3.0)
    assert np.abs(icp_pose - want_icp).max() < 1e-9
    e_r, e_i = synth.pose_errors(ransac_pose, s["T_gt"]), synth.pose_errors(icp_pose, s["T_gt"])
    assert e_r[0] < 0.05 and e_i[0] < 0.05 and e_i[1] < 0.1   # both within the 2 cm point noise of the planted pose
    # the shim's map object: cumulative add_points == one add_points, (N, 3) and (N, 3 + D) stores are separate
    a, b = compat.VoxelHashMap(1.0, 100.0, 20), compat.VoxelHashMap(1.0, 100.0, 20)
    a.add_points(vmap_arr[:, :3])
    b.add_points(vmap_arr[:2000, :3])
    b.add_points(vmap_arr[2000:, :3])
    assert np.array_equal(a.point_cloud(), b.point_cloud()) and np.array_equal(a.point_cloud(), kept_xyz) and a.empty_n()
    posed = s["scan_xyz"].astype(np.float64) @ s["T_gt"][:3, :3].T + s["T_gt"][:3, 3]
    src, tgt = a.get_correspondences(posed, 0.5)
    osrc, otgt = om.get_correspondences(posed, 0.5)
    assert np.array_equal(src, osrc) and np.array_equal(tgt, otgt) and len(src) > 500


@pytest.mark.parametrize("k,outliers", [(400, 40), (37, 0), (1500, 300), (6, 0), (0, 0)])
def test_register_frame_vfm_vs_oracle(vfm, k, outliers):
    """VFM-ICP (Registration.cpp:197-382) given its descriptor correspondences: pruning loop + vanilla loop vs the oracle;
    the surviving count and both iteration counts are equal, the pose agrees to 1e-9 (float64, fixed summation order).
    (Fewer than 3 correspondences make J^T W J singular: the reference's ldlt, NumPy's solve and this LDL^T then return
    different arbitrary updates, so that case is not a parity case.)"""
    rng = np.random.default_rng(100 + k)
    map_pts = _cloud(40 + k, 9000, span=20.0)
    ang = np.deg2rad(4.0)
    T = np.eye(4)
    T[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]
    T[:3, 3] = [0.8, -0.4, 0.05]
    sel = rng.choice(len(map_pts), 2500, replace=False)
    frame = (map_pts[sel] - T[:3, 3]) @ T[:3, :3] + rng.normal(0, 0.01, (2500, 3))
    vs, vt = frame[:k].copy(), map_pts[sel[:k]].copy()
    vs[:outliers] += rng.normal(0, 4.0, (outliers, 3))
    m = vfm.VoxelMap(1.0, 20)
    m.build(map_pts)
    o = ov.VoxelHashMapOracle(1.0, 20)
    o.add_points(map_pts)
    guess = np.eye(4)
    guess[:3, 3] = [0.1, 0.1, 0.0]
    got, gi = vfm.register_frame_vfm(frame, m, vs, vt, guess, 6.0, 2.0 / 3.0, return_info=True)
    want, wi = ov.register_frame_vfm(frame, o, vs, vt, guess, 6.0, 2.0 / 3.0, return_info=True)
    assert gi == wi, (gi, wi)
    assert np.abs(got - want).max() < 1e-9
    if k >= 37:
        assert np.abs(got - T).max() < 0.02
    T1, j, _ = ov.vfm_icp_first_loop(vs, vt, guess, 2.0 / 3.0, 3)
    only1, i1 = vfm.register_frame_vfm(frame[:0], m, vs, vt, guess, 6.0, 2.0 / 3.0, max_iterations=3, return_info=True)
    assert i1["vfm_iterations"] == j and np.abs(only1 - T1).max() < 1e-10


def test_compat_register_frame_with_descriptors(vfm):
    """register_frame on (N, 3 + D) frames = VFM-ICP through the shim (kiss_icp/registration.py:41-62)."""
    from oracle import match
    from vfm_registration_b200 import compat, synth
    s = synth.make_pair(31, 8000, 4000, 32, inlier_frac=0.6)
    vmap_arr = np.c_[s["map_xyz"], s["map_feat"]].astype(np.float32)
    scan_arr = np.c_[s["scan_xyz"], s["scan_feat"]].astype(np.float32)
    guess = s["T_gt"].copy()
    guess[:3, 3] += [0.4, -0.3, 0.05]
    vm = compat.VoxelHashMap(1.0, 100.0, 20)
    vm.add_points(vmap_arr)
    pose = compat.register_frame(scan_arr, vm, guess, 6.0, 2.0 / 3.0)
    # oracle chain
    src_t = scan_arr[:, :3].astype(np.float64) @ guess[:3, :3].T + guess[:3, 3]
    posed = np.c_[src_t, scan_arr[:, 3:]].astype(np.float32)
    vox, idx = ov.voxel_down_sample(posed, 5.0, return_index=True)
    assert len(vox) >= 100
    om = ov.VoxelHashMapOracle(1.0, 20)
    om.add_points(vmap_arr[:, :3])
    kept_xyz, kept_id = om.point_cloud()
    osrc, otgt = match.get_vfm_correspondences(vox, vmap_arr[kept_id], 0.8)
    assert len(osrc) > 50
    mm = match.match_nn(vox[:, 3:], vmap_arr[kept_id, 3:])
    corr = match.filter_correspondences(mm["idx01"], mm["sim01"], min_cos=0.8)
    want = ov.register_frame_vfm(scan_arr[:, :3].astype(np.float64), om, scan_arr[idx[corr[:, 0]], :3].astype(np.float64),
                                 kept_xyz[corr[:, 1]], guess, 6.0, 2.0 / 3.0)
    assert np.abs(pose - want).max() < 1e-8
    rte, rre = synth.pose_errors(pose, s["T_gt"])
    assert rte < 0.05 and rre < 0.1
    assert np.array_equal(compat.register_frame(scan_arr, compat.VoxelHashMap(1.0, 100.0, 20), guess, 6.0, 0.5), guess)


def test_build_local_map_and_prepare_scan_vs_oracle(vfm):
    """registration_node.py:557-590 (local map accumulation, scan thinning) through scenes.py against the oracle chain;
    both return rows in input order, so the arrays are equal element for element."""
    from vfm_registration_b200 import scenes
    rng = np.random.default_rng(12)
    poses, clouds = [], []
    for i in range(4):
        a = 0.2 * i
        T = np.eye(4)
        T[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
        T[:3, 3] = [1.5 * i, -0.7 * i, 0.05 * i]
        pts = _cloud(60 + i, 6000, span=8.0, dtype=np.float32)
        feat = np.abs(rng.standard_normal((6000, 16))).astype(np.float32)
        feat[rng.random(6000) < 0.3] = 0                       # points no camera saw: dropped before the thinning
        feat[5] = [1.0] + [-1.0] + [0.0] * 14                  # sum == 0 but non-zero: dropped by the reference's sum test
        poses.append(T)
        clouds.append(np.c_[pts, feat])
    got = scenes.build_local_map(poses, clouds, voxel_size=0.25, feat_dim=16)
    want = ov.build_local_map(poses, clouds, voxel_size=0.25, feat_dim=16)
    assert got.dtype == np.float32 and got.shape == want.shape and np.array_equal(got, want) and 5000 < len(got) < 17000
    split = scenes.build_local_map(poses, clouds, voxel_size=0.25, feat_dim=16, split_above=1000)
    assert np.array_equal(split, ov.build_local_map(poses, clouds, voxel_size=0.25, feat_dim=16, split_above=1000))
    by_norm = scenes.build_local_map(poses, clouds, voxel_size=0.25, feat_dim=16, has_descriptor="norm")
    assert len(by_norm) >= len(got)
    scan = clouds[0][:, :3]
    assert np.array_equal(scenes.prepare_scan(scan, 0.1), ov.voxel_down_sample(scan, 0.1))


def test_compat_icp_registration(vfm):
    """RegistrationNode.icp_registration (registration_node.py:358-394) for (N, 3) clouds against the oracle chain."""
    from vfm_registration_b200 import compat, synth
    s = synth.make_pair(41, 9000, 4000, 8, inlier_frac=0.8)
    guess = s["T_gt"].copy()
    guess[:3, 3] += [0.3, -0.2, 0.02]
    node = compat.RegistrationNode()
    pose = node.icp_registration(s["map_xyz"], s["scan_xyz"], guess)
    om = ov.VoxelHashMapOracle(1.0, 20)
    om.add_points(s["map_xyz"])
    vs = ov.voxel_down_sample(ov.voxel_down_sample(s["scan_xyz"], 0.5), 1.0)
    want = ov.register_frame(vs.astype(np.float64), om, guess, 6.0, 2.0 / 3.0)
    assert np.abs(pose - want).max() < 1e-9
    rte, rre = synth.pose_errors(pose, s["T_gt"])
    assert rte < 0.05 and rre < 0.1
    with pytest.raises(ValueError, match="Invalid shape"):
        node.icp_registration(s["map_xyz"], s["scan_xyz"][:, :2])


def test_voxel_operations_full_size(vfm):
    """NCLT-scale inputs against the vectorised oracle restatements: 1 M points, a 200k x (3 + 384) cloud, a 200k-point map with
    crowded voxels, and nearest neighbours checked against a KD-tree wherever the true neighbour is closer than one voxel."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(77)
    big = (rng.uniform(-60, 60, (1_000_000, 3)) * np.array([1.0, 1.0, 0.1])).astype(np.float32)
    for vs in (0.25, 1.0):
        _, gi = vfm.voxel_down_sample(big, vs, return_index=True)
        assert np.array_equal(gi, ov.voxel_down_sample_index_fast(big, vs))
    wide = np.c_[big[:200_000], rng.standard_normal((200_000, 384)).astype(np.float32)]
    out, gi = vfm.voxel_down_sample(wide, 0.5, return_index=True)
    wi = ov.voxel_down_sample_index_fast(wide, 0.5)
    assert np.array_equal(gi, wi) and np.array_equal(out, wide[wi])
    dense = np.concatenate([big[:180_000].astype(np.float64), rng.normal(0, 0.3, (20_000, 3))], axis=0)   # ~10^4 points in a few voxels
    m = vfm.VoxelMap(1.0, 20)
    m.build(dense)
    xyz, idx = m.points()
    want = ov.voxel_map_kept_index_fast(dense, 1.0, 20)
    assert np.array_equal(idx.cpu().numpy(), want) and np.array_equal(xyz.cpu().numpy(), dense[want])
    kept = dense[want]
    q = kept[rng.integers(0, len(kept), 20_000)] + rng.normal(0, 0.15, (20_000, 3))
    tgt, valid, d2 = m.nearest(q, 1.0)
    d, j = cKDTree(kept).query(q)
    sure = d < 0.99                                    # the true neighbour then lies inside the 27 voxels
    assert sure.mean() > 0.9 and valid.cpu().numpy()[sure].all()
    assert np.array_equal(tgt.cpu().numpy()[sure], kept[j[sure]]) and np.allclose(np.sqrt(d2.cpu().numpy()[sure]), d[sure], rtol=0, atol=1e-12)


def test_non_finite_points_are_dropped_not_fatal(vfm):
    """Raw scans carry NaN / inf returns: voxel_down_sample and VoxelMap.build drop them (with a warning) and keep reporting
    row numbers of the caller's array (the library alone rejects such a cloud: VFMREG_ERR_ARG)."""
    rng = np.random.default_rng(12)
    pts = rng.uniform(-20, 20, (5000, 3))
    bad = rng.choice(5000, 37, replace=False)
    dirty = pts.copy()
    dirty[bad[:20]] = np.nan
    dirty[bad[20:30], 1] = np.inf
    dirty[bad[30:]] = 1e9   # beyond +-2^20 voxels of 0.5 m
    with pytest.warns(RuntimeWarning, match="37 of 5000"):
        out, idx = vfm.voxel_down_sample(dirty, 0.5, return_index=True)
    good = np.setdiff1d(np.arange(5000), bad)
    ref, ref_idx = vfm.voxel_down_sample(pts[good], 0.5, return_index=True)
    assert np.array_equal(idx, good[ref_idx]) and np.array_equal(out, ref)
    m = vfm.VoxelMap(0.5, 20)
    with pytest.warns(RuntimeWarning):
        m.build(dirty)
    xyz, src = m.points()
    assert np.isfinite(xyz.cpu().numpy()).all() and np.isin(src.cpu().numpy(), good).all()
    assert np.array_equal(xyz.cpu().numpy(), dirty[src.cpu().numpy()])
    m.close()


def test_prepare_scene_producer_on_a_synthetic_nclt_tree(vfm, tmp_path, monkeypatch):
    """scenes.prepare_scene (prepare_scenes.py:110-167) end to end on the miniature NCLT tree: scan decoding, 0.2 / 0.1 m voxel
    thinning, undistorted images, ViT descriptors through the fused projection + gather kernel, scene file.  The descriptors of
    a frame equal a direct create_descriptors call; unseen points carry zero rows; the file round-trips."""
    import json
    import os
    import sys
    pytest.importorskip("cv2")
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import synth_dataset
    import test_scenes_h5_cpu as h5stub
    import types
    fake = types.ModuleType("h5py")
    fake.File = h5stub._File
    monkeypatch.setitem(sys.modules, "h5py", fake)
    from vfm_registration_b200 import datasets, features, scenes
    root = tmp_path / "nclt"
    synth_dataset.build(root, distinct_maps=1)
    ts, seq = synth_dataset.TS, synth_dataset.SEQ
    imgs = [f"images/{seq}/lb3/Cam{c}/{ts}.tiff" for c in range(1, 6)]
    pose = np.eye(4).tolist()
    scene = {"mapping": {"point_clouds": [f"velodyne_data/{seq}/velodyne_sync/{ts}.bin"] * 2, "images": [imgs, imgs], "poses": [pose, pose]},
             "registration": [{"point_cloud": f"velodyne_data/{seq}/velodyne_sync/{ts}.bin", "images": imgs, "pose": pose}]}
    (tmp_path / "scene_000.json").write_text(json.dumps(scene))
    with pytest.warns(RuntimeWarning):
        gen = vfm.ImageFeatureGenerator("dinov2", seed=2, random_init=True)
    out = scenes.prepare_scene(root, tmp_path / "scene_000.json", tmp_path / "processed", gen, dataset="nclt")
    s = scenes.read_scenes(out)
    assert len(s["map_point_clouds"]) == 2 and len(s["scene_point_clouds"]) == 1
    m0, sc = s["map_point_clouds"][0], s["scene_point_clouds"][0]
    assert m0.shape[1] == 3 + 384 and sc.shape[1] == 387 and sc.shape[0] >= m0.shape[0]     # 0.1 m keeps at least the points 0.2 m keeps
    loader = datasets.NCLT(seq, root)
    pcl = vfm.voxel_down_sample(loader.read_pcl(frame_id=0), 0.2).astype(np.float32)
    assert np.array_equal(m0[:, :3], pcl)
    images = loader.read_images(frame_id=0)
    want = features.create_descriptors(images, loader.project_params(images), gen, pcl)
    assert np.array_equal(m0[:, 3:], want)
    seen = np.abs(m0[:, 3:]).sum(axis=1) > 0
    assert 0 < (~seen).sum() < len(seen)      # some points project into no camera: zero rows (prepare_scenes.py:102-104)
