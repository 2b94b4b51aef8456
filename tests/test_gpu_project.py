"""GPU parity of the fused projection + gather kernel against the oracle restatements of project_pcl_to_image /
create_descriptors (themselves pinned to the reference's functions by tests/golden)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import project  # noqa: E402


@pytest.fixture(scope="module")
def vfm():
    assert torch.cuda.is_available()
    import vfm_registration_b200 as v
    v.get_context(0)
    return v


def _scene(rng, n=5000, n_cam=3, hw=(70, 82), grid=(16, 18), d=384):
    from scipy.spatial.transform import Rotation as R
    pts = np.c_[rng.uniform(-15, 15, (n, 2)), rng.uniform(-2, 4, n)].astype(np.float32)
    images = [rng.integers(0, 255, (*hw, 3), dtype=np.uint8) for _ in range(n_cam)]
    images[0][5:25, 10:40] = 0
    images[-1][:, :9] = 0
    toks = [rng.standard_normal((*grid, d)).astype(np.float32) for _ in range(n_cam)]
    ks, ts = [], []
    for i in range(n_cam):
        ks.append(np.array([[40.0 + 3 * i, 0, hw[1] / 2], [0, 41.0 - i, hw[0] / 2], [0, 0, 1.0]]))
        t = np.eye(4)
        t[:3, :3] = (R.from_euler("z", 50.0 * i, degrees=True) * R.from_euler("yx", [90, -90], degrees=True)).as_matrix().T
        t[:3, 3] = [0.1 * i, 0.0, 0.2]
        ts.append(t)
    return pts, images, toks, ks, ts


def _oracle(pts, images, toks, ks, ts, hw, black=True):
    feats = {i: np.ascontiguousarray(project.upsample_bilinear(np.ascontiguousarray(t.transpose(2, 0, 1)), *hw).transpose(1, 2, 0))
             for i, t in enumerate(toks)}
    imgs = {i: im for i, im in enumerate(images)}

    def fn(pcl_h, image, cam):
        return project.project_pinhole(pcl_h[:3].T, ks[cam], ts[cam], hw[0], hw[1], image=image, reject_black=black)
    return project.create_descriptors(imgs, feats, fn, pts)


@pytest.mark.parametrize("d", [384, 100, 6])
def test_project_gather_vs_create_descriptors(vfm, d):
    rng = np.random.default_rng(d)
    hw, grid = (70, 82), (16, 18)
    pts, images, toks, ks, ts = _scene(rng, d=d, hw=hw, grid=grid)
    cams = [vfm.CameraSpec(P=ks[i] @ ts[i][:3], img_hw=hw, grid_hw=grid, black_mode=1) for i in range(len(ks))]
    desc, cam_of, uv = vfm.project_gather(pts, cams, toks, images)
    want = _oracle(pts, images, toks, ks, ts, hw)
    got = desc.cpu().numpy()
    seen = np.abs(want).sum(1) > 0
    assert 0.2 < seen.mean() < 0.95
    assert np.array_equal(cam_of.cpu().numpy() >= 0, seen)          # same visible set (bit-exact pixel decisions)
    assert np.abs(got - want).max() < 1e-5                           # fp32 bilinear blend tolerance
    assert np.all(got[~seen] == 0)
    # per-camera pixel indices equal the oracle's projection (first camera wins)
    taken = np.zeros(len(pts), dtype=bool)
    for c in range(len(ks)):
        u, v, idx = project.project_pinhole(pts, ks[c], ts[c], hw[0], hw[1], image=images[c])
        new = idx[~taken[idx]]
        sel = cam_of.cpu().numpy() == c
        assert np.array_equal(np.nonzero(sel)[0], np.sort(new))
        lut = {int(i): (int(a), int(b)) for i, a, b in zip(idx, u, v)}
        g = uv.cpu().numpy()[sel]
        assert all(lut[int(i)] == (int(a), int(b)) for i, (a, b) in zip(np.nonzero(sel)[0], g))
        taken[idx] = True


def test_project_gather_golden_nclt(vfm, golden):
    """The NCLT recipe (crop window, sub-sampling, truncation before the window test, black-pixel rejection) on the
    reference-generated fixture: visible set and pixel coordinates are bit-exact."""
    g = golden("project_nclt.npz")
    pts = g["pts"].astype(np.float32)
    sub = int(g["sub"])
    mc = g["coords"] // sub
    img = g["image"]
    tok = np.random.default_rng(0).standard_normal((4, 5, 8)).astype(np.float32)
    cam = vfm.CameraSpec(P=g["k"] @ g["t_c_body"][:3], img_hw=img.shape[:2], grid_hw=(4, 5), crop=(mc[0], mc[1], mc[2], mc[3]),
                         subsample=sub, black_mode=1)
    desc, cam_of, uv = vfm.project_gather(pts, [cam], [tok], [img])
    sel = np.nonzero(cam_of.cpu().numpy() == 0)[0]
    assert np.array_equal(sel, g["idx"])
    assert np.array_equal(uv.cpu().numpy()[sel, 0], g["x_im"]) and np.array_equal(uv.cpu().numpy()[sel, 1], g["y_im"])


def test_project_gather_golden_oxford(vfm, golden):
    g = golden("project_oxford.npz")
    pts = g["pts"].astype(np.float32)
    img = g["image"]
    sub = int(g["sub"])
    m = np.linalg.solve(g["g"], g["cam_in_ego"] @ g["lidar_in_ego"])
    kmat = np.array([[g["focal"][0], 0, g["principal"][0]], [0, g["focal"][1], g["principal"][1]], [0, 0, 1.0]])
    tok = np.random.default_rng(1).standard_normal((3, 4, 8)).astype(np.float32)
    cam = vfm.CameraSpec(P=kmat @ m[:3], img_hw=img.shape[:2], grid_hw=(3, 4), subsample=sub, z_inclusive=True,
                         float_bounds=True, black_mode=0)
    desc, cam_of, uv = vfm.project_gather(pts, [cam], [tok], None)
    sel = np.nonzero(cam_of.cpu().numpy() == 0)[0]
    assert np.array_equal(sel, g["idx"])
    assert np.array_equal(uv.cpu().numpy()[sel, 0], g["u"]) and np.array_equal(uv.cpu().numpy()[sel, 1], g["v"])


def test_project_gather_rot90_matches_reference_nclt_branch(vfm, golden):
    """prepare_scenes.py's NCLT branch (rot90 of image and features around the projection) on the golden fixture."""
    g = golden("create_descriptors.npz")
    pts, ks, ts = g["pts"], g["ks"], g["ts"]
    images, feats = g["images2"], g["feats2"]           # square (48, 48), full-resolution features
    hw = images.shape[1:3]
    # a token grid equal to the full-resolution map makes the bilinear sampler an identity lookup
    cams = [vfm.CameraSpec(P=ks[i] @ ts[i][:3], img_hw=hw, grid_hw=hw, black_mode=1, rot90=True) for i in range(3)]
    zeroed = feats.copy()
    zeroed[np.all(images == 0, axis=-1)] = 0
    desc, cam_of, _ = vfm.project_gather(pts, cams, [zeroed[i] for i in range(3)], [images[i] for i in range(3)])
    assert np.abs(desc.cpu().numpy() - g["out2"]).max() < 1e-6
    cams0 = [vfm.CameraSpec(P=ks[i] @ ts[i][:3], img_hw=g["images"].shape[1:3], grid_hw=g["images"].shape[1:3], black_mode=1)
             for i in range(3)]
    z0 = g["feats"].copy()
    z0[np.all(g["images"] == 0, axis=-1)] = 0
    d0, _, _ = vfm.project_gather(pts, cams0, [z0[i] for i in range(3)], [g["images"][i] for i in range(3)])
    assert np.abs(d0.cpu().numpy() - g["out"]).max() < 1e-6


def test_nclt_dataset_camera_spec_on_the_kernel(vfm, golden, tmp_path):
    """datasets.NCLT.camera_spec drives the fused kernel to the visible set and pixels of the reference's
    project_pcl_to_image on the synthetic NCLT tree (fixture from oracle/gen_golden_datasets.py)."""
    import os
    import sys
    pytest.importorskip("cv2")
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import synth_dataset
    from vfm_registration_b200 import datasets
    g = golden("datasets_nclt.npz")
    synth_dataset.build(tmp_path, cameras=("Cam1",), distinct_maps=1)
    for sub in (1, 2):
        seq = datasets.NCLT(synth_dataset.SEQ, tmp_path, image_subsample=sub, cameras=("Cam1",))
        img = seq.read_images(frame_id=0)["Cam1"]
        pcl = seq.read_pcl(frame_id=0)
        spec = seq.camera_spec("Cam1", img.shape[:2])
        spec.grid_hw = (4, 5)
        tok = np.random.default_rng(0).standard_normal((4, 5, 8)).astype(np.float32)
        _, cam_of, uv = vfm.project_gather(pcl, [spec], [tok], [img])
        sel = np.nonzero(cam_of.cpu().numpy() == 0)[0]
        assert np.array_equal(sel, g[f"proj{sub}_Cam1_idx"])
        assert np.array_equal(uv.cpu().numpy()[sel, 0], g[f"proj{sub}_Cam1_x"])
        assert np.array_equal(uv.cpu().numpy()[sel, 1], g[f"proj{sub}_Cam1_y"])


def test_binned_path_equals_one_kernel_path(vfm):
    """Clouds of >= 65536 points are binned by (camera, token cell) before the gather (project.cu); the result must be the
    one-kernel path's bit for bit -- the same cloud in chunks below the threshold runs that path -- including the rot90 frame,
    black pixels that hide (mode 1) or claim (mode 2) a point, and points no camera sees."""
    rng = np.random.default_rng(21)
    n, d, hw, grid = 200_000, 384, (96, 128), (7, 9)
    pts = np.c_[rng.uniform(-30, 30, (n, 2)), rng.uniform(-3, 6, n)].astype(np.float32)
    kmat = np.array([[60.0, 0, 64.0], [0, 60.0, 48.0], [0, 0, 1.0]])
    cams, toks, imgs = [], [], []
    from scipy.spatial.transform import Rotation as R
    for c in range(4):
        t = np.eye(4)
        t[:3, :3] = (R.from_euler("z", 90.0 * c, degrees=True) * R.from_euler("yx", [90, -90], degrees=True)).as_matrix().T
        rot = c == 1
        img = rng.integers(1, 255, ((hw[1], hw[0]) if rot else hw) + (3,), dtype=np.uint8)
        img[10:30, 20:50] = 0
        cams.append(vfm.CameraSpec(P=kmat @ t[:3], img_hw=hw, grid_hw=grid if not rot else (grid[1], grid[0]), black_mode=(1, 1, 2, 0)[c], rot90=rot))
        toks.append(rng.standard_normal(((grid[1], grid[0]) if rot else grid) + (d,)).astype(np.float32))
        imgs.append(img)
    desc, cam_of, uv = vfm.project_gather(pts, cams, toks, imgs)
    desc, cam_of, uv = desc.cpu().numpy(), cam_of.cpu().numpy(), uv.cpu().numpy()
    assert (cam_of >= 0).mean() > 0.3 and (cam_of < 0).mean() > 0.01 and len(np.unique(cam_of)) == 5
    for lo in range(0, n, 50_000):
        d2, c2, u2 = vfm.project_gather(pts[lo:lo + 50_000], cams, toks, imgs)
        assert np.array_equal(cam_of[lo:lo + 50_000], c2.cpu().numpy()) and np.array_equal(uv[lo:lo + 50_000], u2.cpu().numpy())
        assert np.array_equal(desc[lo:lo + 50_000], d2.cpu().numpy())
