"""A miniature NCLT-format tree (one timestamp) used to pin vfm_registration_b200.datasets to the reference's loader:
oracle/gen_golden_datasets.py runs /root/reference's NCLT class on it and stores the outputs; tests/test_datasets_cpu.py
rebuilds the identical tree (seeded) and runs the product loader."""
import os
from pathlib import Path

import numpy as np

SEQ, TS = "2012-01-08", 1326030975726043
W, H = 1616, 1232


def distortion_maps(cam_index: int):
    """Smooth synthetic undistortion maps (mapu = source column, mapv = source row of each undistorted pixel)."""
    r, c = np.mgrid[0:H, 0:W].astype(np.float64)
    dx, dy = (c - W / 2) / W, (r - H / 2) / H
    rad = dx * dx + dy * dy
    k = 0.08 + 0.01 * cam_index
    return c + k * W * dx * rad + 3.0 * np.sin(r / 97.0), r + k * H * dy * rad + 2.0 * np.cos(c / 131.0)


def write_u2d(path: Path, cam_index: int):
    mapu, mapv = distortion_maps(cam_index)
    r, c = np.mgrid[0:H, 0:W]
    rows = np.stack([r.ravel(), c.ravel(), mapv.ravel(), mapu.ravel()], axis=1)
    with open(path, "w") as f:
        f.write(f"{W}, {H}\n")
        np.savetxt(f, rows, fmt="%d %d %.4f %.4f")


def build(root, cameras=("Cam1", "Cam2", "Cam3", "Cam4", "Cam5"), distinct_maps=2):
    """Writes the tree under `root`; cameras beyond `distinct_maps` re-use (hard-link) an earlier camera's 60 MB map file."""
    import cv2
    root = Path(root)
    (root / "cam_params").mkdir(parents=True, exist_ok=True)
    for cam in cameras:
        n = cam[-1]
        i = int(n) - 1   # every parameter depends on the camera number only: a subset of cameras builds the same files
        p = root / "cam_params" / f"U2D_{cam}_{W}X{H}.txt"
        if not p.exists():
            src = root / "cam_params" / f"U2D_Cam{i % distinct_maps + 1}_{W}X{H}.txt"
            if not src.exists():
                write_u2d(src, i % distinct_maps)
            if src != p:
                os.link(src, p)
        K = np.array([[400.0 + 5 * i, 0, 810.0 + i], [0, 402.0 + 3 * i, 615.0 - i], [0, 0, 1]])
        np.savetxt(root / "cam_params" / f"K_cam{n}.csv", K, delimiter=",")
        x = np.array([0.04 * np.cos(1.2566 * i), 0.04 * np.sin(1.2566 * i), 0.0, 0.5 * i - 90.0, 0.3 * i, 72.0 * i - 10.0])
        np.savetxt(root / "cam_params" / f"x_lb3_c{n}.csv", x[None], delimiter=",")
        d = root / "images" / SEQ / "lb3" / cam
        d.mkdir(parents=True, exist_ok=True)
        rng = np.random.default_rng(1000 + int(n))   # per camera: a subset of cameras builds the same files
        img = cv2.resize(rng.integers(0, 256, (H // 8, W // 8, 3), dtype=np.uint8), (W, H), interpolation=cv2.INTER_LINEAR)
        img[300:340, 500:900] = 0   # a black band inside the crop window (black-pixel rejection)
        cv2.imwrite(str(d / f"{TS}.tiff"), img)
    v = root / "velodyne_data" / SEQ / "velodyne_sync"
    v.mkdir(parents=True, exist_ok=True)
    n_pts = 6000
    rng = np.random.default_rng(20260117)
    xyz = rng.uniform(-60, 60, (n_pts, 3)) * np.array([1, 1, 0.1])
    raw = np.zeros((n_pts, 4), dtype=np.int16)
    raw[:, :3] = np.round((xyz + 100.0) / 0.005).astype(np.int16)
    raw[:, 3] = rng.integers(0, 255, n_pts)
    raw.tofile(v / f"{TS}.bin")
    return root


# ---- RobotCar: a camera model (intrinsics + distortion look-up table) and a Bayer image -------------------------------------
def write_robotcar_models(models_dir, h, w, names=("mono_left",)):
    """<name>.txt (fx fy cx cy + G_camera_image) and <name>_distortion_lut.bin (float64 (2, H*W): u then v of the source pixel)."""
    models_dir = Path(models_dir)
    v, u = np.mgrid[0:h, 0:w].astype(np.float64)
    du, dv = (u - w / 2) / w, (v - h / 2) / h
    r2 = du * du + dv * dv
    lut = np.stack([(u + 0.12 * w * du * r2 + 1.5 * np.sin(v / 37.0)).ravel(), (v + 0.12 * h * dv * r2 + 1.1 * np.cos(u / 41.0)).ravel()])
    for n in names:
        (models_dir / f"{n}.txt").write_text("400.0 401.0 %.1f %.1f\n0 0 1 0\n1 0 0 0\n0 1 0 0\n0 0 0 1\n" % (w / 2, h / 2))
        lut.tofile(models_dir / f"{n}_distortion_lut.bin")


def robotcar_cfa(h, w):
    rng = np.random.default_rng(77)
    import cv2
    return cv2.resize(rng.integers(0, 256, (h // 4, w // 4), dtype=np.uint8), (w, h), interpolation=cv2.INTER_LINEAR)
