"""vfm_registration_b200.datasets against the reference's own NCLT loader: tests/golden/datasets_nclt.npz was produced by
oracle/gen_golden_datasets.py running /root/reference's dataloader/nclt.py on the tree tests/synth_dataset.py builds."""
import hashlib
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth_dataset  # noqa: E402

cv2 = pytest.importorskip("cv2")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "datasets_nclt.npz")


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    root = tmp_path_factory.mktemp("nclt")
    synth_dataset.build(root, cameras=("Cam1", "Cam2", "Cam5"))
    return root


def test_nclt_loader_matches_reference(tree):
    from vfm_registration_b200 import datasets
    g = np.load(GOLD)
    seq = datasets.NCLT(synth_dataset.SEQ, tree, cameras=("Cam1", "Cam2", "Cam5"))
    assert seq.read_times()["pcl"] == g["timestamps"].tolist() and len(seq) == 1 and seq.timestamps == [0.0]
    pcl = seq.read_pcl(frame_id=0)
    assert pcl.dtype == np.float32 and np.array_equal(pcl, g["pcl"])          # decode + 50 m crop, bit for bit
    assert np.array_equal(seq.read_pcl(filename=seq.pcl_file(0)), g["pcl"])
    np.testing.assert_allclose(seq.calib["lidar_in_ego"], g["lidar_in_ego"], rtol=0, atol=1e-15)
    for cam in ("Cam1", "Cam2", "Cam5"):
        p = seq.read_camera_parameters(cam)
        assert np.array_equal(p["K"], g[f"K_{cam}"])
        np.testing.assert_allclose(p["x_lb3"], g[f"x_lb3_{cam}"], rtol=0, atol=1e-15)
    for sub in (1, 2):
        s = datasets.NCLT(synth_dataset.SEQ, tree, image_subsample=sub, cameras=("Cam1", "Cam2"))
        s._maps = seq._maps   # the parsed maps are per camera, not per sub-sampling
        imgs = s.read_images(frame_id=0)
        for cam in ("Cam1", "Cam2"):
            im = imgs[cam]
            assert list(im.shape) == g[f"img{sub}_{cam}_shape"].tolist()
            assert np.array_equal(im[100:116, 200:216], g[f"img{sub}_{cam}_patch"])
            sha = np.frombuffer(hashlib.sha256(np.ascontiguousarray(im).tobytes()).digest(), dtype=np.uint8)
            assert np.array_equal(sha, g[f"img{sub}_{cam}_sha"]), "one fused remap must equal remap -> crop -> rotate, pixel for pixel"


def test_nclt_camera_spec_reproduces_reference_projection(tree):
    """The CameraSpec handed to the fused GPU kernel, evaluated by the CPU oracle of that kernel (oracle/project.py), selects
    the same points and pixels as the reference's project_pcl_to_image on the same scan and image."""
    from oracle import project
    from vfm_registration_b200 import datasets
    g = np.load(GOLD)
    for sub in (1, 2):
        seq = datasets.NCLT(synth_dataset.SEQ, tree, image_subsample=sub, cameras=("Cam1", "Cam2"))
        imgs = seq.read_images(frame_id=0)
        pcl = seq.read_pcl(frame_id=0)
        for cam in ("Cam1", "Cam2"):
            spec = seq.camera_spec(cam, imgs[cam].shape[:2])
            assert spec.rot90 and spec.img_hw == (imgs[cam].shape[1], imgs[cam].shape[0])
            unrot = cv2.rotate(imgs[cam], cv2.ROTATE_90_COUNTERCLOCKWISE)
            K = seq.read_camera_parameters(cam)["K"]
            p4 = np.insert(pcl.astype(np.float64), 3, values=1, axis=1).T
            assert spec.crop == tuple(v // sub for v in datasets.NCLT_CROP) and spec.subsample == sub and spec.black_mode == 1
            np.testing.assert_allclose(spec.P, K @ seq.camera_from_body(cam)[:3], atol=0)
            x, y, idx = project.project_nclt(p4, unrot, K, seq.camera_from_body(cam), sub, datasets.NCLT_CROP)
            assert np.array_equal(idx, g[f"proj{sub}_{cam}_idx"])
            assert np.array_equal(x, g[f"proj{sub}_{cam}_x"]) and np.array_equal(y, g[f"proj{sub}_{cam}_y"])


def test_robotcar_loader(tmp_path):
    from vfm_registration_b200 import datasets
    sdk = tmp_path / "sdk"
    (sdk / "extrinsics").mkdir(parents=True)
    (sdk / "models").mkdir()
    rng = np.random.default_rng(3)
    ext = {n: rng.uniform(-1, 1, 6) for n in ("velodyne_left", "stereo", "mono_left", "mono_right", "mono_rear", "ins")}
    for n, v in ext.items():
        (sdk / "extrinsics" / f"{n}.txt").write_text(" ".join(repr(float(x)) for x in v) + "\n")
    G = np.array([[0, 0, 1, 0], [1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1.0]])
    for n in ("stereo_narrow_left", "mono_left", "mono_right", "mono_rear"):
        (sdk / "models" / f"{n}.txt").write_text("400.0 401.0 500.5 480.25\n" + "\n".join(" ".join(str(x) for x in r) for r in G) + "\n")
    seq = datasets.OxfordRobotcar("2019-01-10-11-46-21", tmp_path, sdk)
    # build_se3_transform: R = Rz(yaw) Ry(pitch) Rx(roll) (robotcar_sdk/python/transform.py:47-68)
    x, y, z, r, p, w = ext["velodyne_left"]
    Rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
    Ry = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
    Rz = np.array([[np.cos(w), -np.sin(w), 0], [np.sin(w), np.cos(w), 0], [0, 0, 1]])
    np.testing.assert_allclose(seq.calib["lidar_in_ego"][:3, :3], Rz @ Ry @ Rx, atol=1e-15)
    np.testing.assert_allclose(seq.calib["lidar_in_ego"][:3, 3], [x, y, z])
    np.testing.assert_allclose(seq.calib["ins_in_ego"] @ seq.calib["lidar_in_ins"], seq.calib["lidar_in_ego"], atol=1e-12)
    # velodyne_left record: float32 (4, N); ego-vehicle returns (< 2.5 m) and far returns (>= 50 m) dropped
    pts = np.concatenate([rng.uniform(-1, 1, (50, 3)), rng.uniform(-30, 30, (400, 3)), rng.uniform(60, 70, (20, 3))]).astype(np.float32)
    raw = np.concatenate([pts, rng.uniform(0, 1, (len(pts), 1)).astype(np.float32)], axis=1)
    f = tmp_path / "scan.bin"
    np.ascontiguousarray(raw.T).tofile(f)
    out = seq.read_pcl(filename=f)
    d = np.linalg.norm(pts, axis=1)
    assert np.array_equal(out, pts[(d > 2.5) & (d < 50)])
    spec = seq.camera_spec("stereo/centre", (760, 1280))
    assert spec.float_bounds and spec.z_inclusive and spec.black_mode == 2 and spec.P.shape == (3, 4)


def test_robotcar_raw_image_chain(tmp_path):
    """Bayer PNG -> bilinear demosaic -> LUT undistortion -> uint8 -> crop (oxford_robotcar.py:101-137).  The undistortion is
    pinned to the reference's own CameraModel (robotcar_sdk/python/camera_model.py, fixture from oracle/gen_golden_datasets.py);
    the demosaic (colour_demosaicing is not installable offline) is checked against direct neighbour averaging."""
    from PIL import Image
    from vfm_registration_b200 import datasets
    g = np.load(GOLD)
    h, w = int(g["rc_hw"][0]), int(g["rc_hw"][1])
    sdk = tmp_path / "sdk"
    (sdk / "extrinsics").mkdir(parents=True)
    (sdk / "models").mkdir()
    for n in ("velodyne_left", "stereo", "mono_left", "mono_right", "mono_rear", "ins"):
        (sdk / "extrinsics" / f"{n}.txt").write_text("0 0 0 0 0 0\n")
    synth_dataset.write_robotcar_models(sdk / "models", h, w)
    cfa8 = synth_dataset.robotcar_cfa(h, w)
    cfa = cfa8.astype(np.float64)
    seq = datasets.OxfordRobotcar("2019-01-10-11-46-21", tmp_path, sdk, cameras=("mono_left",))
    rgb = datasets.demosaic_bilinear(cfa, "RGGB")
    # interior pixels: red at (even, even), green at the two mixed sites, blue at (odd, odd)
    y, x = 10, 12                                         # a red site
    assert rgb[y, x, 0] == cfa[y, x]
    assert rgb[y, x, 1] == (cfa[y - 1, x] + cfa[y + 1, x] + cfa[y, x - 1] + cfa[y, x + 1]) / 4
    assert rgb[y, x, 2] == (cfa[y - 1, x - 1] + cfa[y - 1, x + 1] + cfa[y + 1, x - 1] + cfa[y + 1, x + 1]) / 4
    y, x = 11, 12                                         # a green site on a blue row
    assert rgb[y, x, 1] == cfa[y, x] and rgb[y, x, 0] == (cfa[y - 1, x] + cfa[y + 1, x]) / 2 and rgb[y, x, 2] == (cfa[y, x - 1] + cfa[y, x + 1]) / 2
    und = seq.undistort("mono_left", rgb)
    # the reference's CameraModel.undistort on the same demosaiced image: every value (sha-256) and a patch
    sha = np.frombuffer(hashlib.sha256(np.ascontiguousarray(und).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(und[40:56, 60:76], g["rc_undistorted_patch"]) and np.array_equal(sha, g["rc_undistorted_sha"])
    Image.fromarray(cfa8).save(tmp_path / "raw.png")
    out = seq.read_images([tmp_path / "raw.png"], raw=True)["mono_left"]
    assert out.shape == (h - 200, w, 3) and out.dtype == np.uint8
    assert np.array_equal(out, und.astype(np.uint8)[: h - 200])
