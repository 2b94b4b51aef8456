"""The NumPy oracle against outputs of the reference's own Python functions
(tests/golden/*.npz, produced by oracle/gen_golden.py from /root/reference)."""
import numpy as np

from oracle import match, metrics, project, ransac


def test_project_nclt(golden):
    g = golden("project_nclt.npz")
    pcl_h = np.insert(g["pts"], 3, values=1, axis=1).T
    x, y, idx = project.project_nclt(pcl_h, g["image"], g["k"], g["t_c_body"], int(g["sub"]), g["coords"])
    assert np.array_equal(idx, g["idx"]) and np.array_equal(x, g["x_im"]) and np.array_equal(y, g["y_im"])
    assert len(idx) > 100


def test_project_oxford(golden):
    g = golden("project_oxford.npz")
    pcl_h = np.insert(g["pts"], 3, values=1, axis=1).T
    u, v, idx = project.project_oxford(pcl_h, g["image"], g["lidar_in_ego"], g["cam_in_ego"], g["g"], g["focal"],
                                       g["principal"], int(g["sub"]))
    assert np.array_equal(idx, g["idx"]) and np.array_equal(u, g["u"]) and np.array_equal(v, g["v"])


def _proj_fn(ks, ts, cams):
    def fn(pcl_h, image, camera):
        i = cams.index(camera)
        return project.project_pinhole(pcl_h[:3].T, ks[i], ts[i], image.shape[0], image.shape[1], image=image)
    return fn


def test_create_descriptors(golden):
    g = golden("create_descriptors.npz")
    cams = ["c0", "c1", "c2"]
    fn = _proj_fn(g["ks"], g["ts"], cams)
    images = {c: g["images"][i] for i, c in enumerate(cams)}
    feats = {c: g["feats"][i] for i, c in enumerate(cams)}
    out = project.create_descriptors(images, feats, fn, g["pts"])
    assert np.array_equal(out, g["out"])
    images2 = {c: g["images2"][i] for i, c in enumerate(cams)}
    feats2 = {c: g["feats2"][i] for i, c in enumerate(cams)}
    out2 = project.create_descriptors(images2, feats2, fn, g["pts"], nclt_rot90=True)
    assert np.array_equal(out2, g["out2"])
    seen = np.abs(out).sum(1) > 0
    assert 100 < seen.sum() < len(seen)


def test_metrics_and_transform(golden):
    g = golden("metrics.npz")
    errs = np.array([metrics.compute_errors(p, q) for p, q in zip(g["poses"], g["gts"])])
    assert np.array_equal(errs, g["errs"])
    rates = [metrics.success_rate(errs[:, 0], errs[:, 1], t, r) for t, r in ((0.3, 15), (0.6, 1.5), (2, 5), (1, 5))]
    assert np.array_equal(np.array(rates), g["rates"])
    assert np.array_equal(metrics.transform_pcl(g["pcl"], g["poses"][3]), g["pcl_t"])


def test_kabsch_vs_pointdsc(golden):
    g = golden("kabsch_pointdsc.npz")
    n_reflect = 0
    for a, b, k, t in zip(g["a"], g["b"], g["k"], g["t"]):
        a32, b32 = a[:k].astype(np.float32).astype(np.float64), b[:k].astype(np.float32).astype(np.float64)
        r, tr, ok = ransac.kabsch(a32, b32)
        assert ok
        assert abs(np.linalg.det(r) - 1) < 1e-9
        # the reference computes in float32 (torch.svd): tolerance is its precision, not ours
        assert np.abs(r - t[:3, :3]).max() < 5e-4 and np.abs(tr - t[:3, 3]).max() < 2e-2
        u, s, vt = np.linalg.svd((b32 - b32.mean(0)).T @ (a32 - a32.mean(0)))
        n_reflect += np.linalg.det(u) * np.linalg.det(vt) < 0
    assert n_reflect >= 4  # the det(U)det(V) = -1 branch is exercised


def test_find_correspondences(golden):
    g = golden("find_corr.npz")
    i0, i1 = match.find_correspondences(g["f0"], g["f1"], mutual_filter=True)
    assert np.array_equal(i0, g["i0"]) and np.array_equal(i1, g["i1"])
    j0, j1 = match.find_correspondences(g["f0"], g["f1"], n_points=100, mutual_filter=False)
    o = np.argsort(j0)
    assert np.array_equal(j0[o], g["j0"]) and np.array_equal(j1[o], g["j1"])
    # the same mutual set through the inner-product formulation used on the GPU (unit vectors)
    f0 = g["f0"] / np.linalg.norm(g["f0"], axis=1, keepdims=True)
    f1 = g["f1"] / np.linalg.norm(g["f1"], axis=1, keepdims=True)
    r = match.match_nn(f0.astype(np.float32), f1.astype(np.float32), normalize=True, mutual=True)
    corr = match.filter_correspondences(r["idx01"], r["sim01"], idx10=r["idx10"], mutual=True)
    k0, k1 = match.find_correspondences(f0, f1, mutual_filter=True)
    assert np.array_equal(corr[:, 0], k0) and np.array_equal(corr[:, 1], k1)


def test_upsample(golden):
    g = golden("upsample.npz")
    up = project.upsample_bilinear(g["feat"], g["up"].shape[1], g["up"].shape[2])
    assert np.abs(up - g["up"]).max() < 1e-5  # float32 evaluation-order tolerance
