"""Raw-data side of the producer: NCLT and Oxford Radar RobotCar sequences as the reference's loaders read
them (dataloader/nclt.py:21-220,290-366, dataloader/oxford_robotcar.py:23-217,330-363) -- file layout, calibration, point
cloud decoding, image undistortion / crop / rotation -- plus what the reference lacks: the per-camera ``CameraSpec`` that lets
``features.create_descriptors`` run projection + gather as one GPU kernel instead of ``project_pcl_to_image``'s Python loops.

Host code by design (file parsing and OpenCV warps; nothing here is on the measured hot path).  Differences from the
reference, none of which changes a pixel:
  * NCLT images: the reference remaps the full 1616 x 1232 frame, copies it through a same-size bicubic resize (an identity),
    converts BGR->RGB, crops the 700 x 820 window and rotates by 90 degrees -- five full-frame passes.  Here the undistortion
    maps are cropped and rotated ONCE when the calibration is read, so one cv2.remap writes the final 820 x 700 image
    (remap is per-pixel, hence the same fixed-point bilinear samples; tests/test_datasets_cpu.py checks equality against the
    reference's own output);
  * the U2D text maps (2 M lines per camera) are parsed with one vectorised pass instead of a Python loop;
  * RobotCar extrinsics / camera models / distortion LUTs are looked up in a directory the caller names (the SDK is not
    vendored here); raw Bayer images go through the same demosaic -> LUT undistortion -> crop chain as the reference's loader.
"""
from __future__ import annotations

import re
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import api

NCLT_CAMERAS = ("Cam1", "Cam2", "Cam3", "Cam4", "Cam5")          # nclt.py:38-39 (Cam0 looks at the sky)
NCLT_CROP = (210, 450, 820, 700)                                   # y0, x0, h, w of the undistorted frame (nclt.py:189)
NCLT_FRAME = (1232, 1616)                                          # nclt.py:193
ROBOTCAR_CAMERAS = ("stereo/centre", "mono_left", "mono_right", "mono_rear")   # oxford_robotcar.py:35-37


def euler_xyz_deg(angles: Sequence[float]) -> np.ndarray:
    """scipy's Rotation.from_euler("xyz", angles, degrees=True).as_matrix(): extrinsic x, y, z = Rz @ Ry @ Rx."""
    rx, ry, rz = np.deg2rad(np.asarray(angles, dtype=np.float64))
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def se3_from_xyzrpy(xyzrpy: Sequence[float]) -> np.ndarray:
    """robotcar_sdk/python/transform.py:23-68 (build_se3_transform): translation + roll/pitch/yaw in radians, R = Rz Ry Rx."""
    if len(xyzrpy) != 6:
        raise ValueError("Must supply 6 values to build transform")
    x, y, z, r, p, w = (float(v) for v in xyzrpy)
    T = np.eye(4)
    T[:3, :3] = euler_xyz_deg(np.rad2deg([r, p, w]))
    T[:3, 3] = (x, y, z)
    return T


def _pose(xyz, rpy_deg) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = euler_xyz_deg(rpy_deg)
    T[:3, 3] = xyz
    return T


def demosaic_bilinear(cfa: np.ndarray, pattern: str) -> np.ndarray:
    """colour_demosaicing.demosaicing_CFA_Bayer_bilinear (the library the reference calls, oxford_robotcar.py:8,108-111; not
    installable offline, restated from its published source): each colour plane is the CFA masked to that colour's sites and
    convolved with the bilinear kernel -- [[1,2,1],[2,4,2],[1,2,1]] / 4 for red and blue, [[0,1,0],[1,4,1],[0,1,0]] / 4 for green --
    with reflected borders (scipy.ndimage.convolve's default).  ``pattern``: the 2 x 2 Bayer tile, row-major ("RGGB", "GBRG")."""
    from scipy.ndimage import convolve
    cfa = np.asarray(cfa, dtype=np.float64)
    masks = {c: np.zeros(cfa.shape, dtype=np.float64) for c in "RGB"}
    for ch, (y, x) in zip(pattern.upper(), ((0, 0), (0, 1), (1, 0), (1, 1))):
        masks[ch][y::2, x::2] = 1.0
    h_g = np.array([[0, 1, 0], [1, 4, 1], [0, 1, 0]], dtype=np.float64) / 4
    h_rb = np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=np.float64) / 4
    return np.stack([convolve(cfa * masks["R"], h_rb), convolve(cfa * masks["G"], h_g), convolve(cfa * masks["B"], h_rb)], axis=2)


class NCLT:
    """One NCLT sequence (reference: dataloader/nclt.py).  ``root_dir`` holds cam_params/, images/<seq>/lb3/CamN/*.tiff and
    velodyne_data/<seq>/velodyne_sync/*.bin.  Calibration is read lazily, so ``read_pcl`` works on a bare velodyne folder."""

    def __init__(self, sequence: str, root_dir, image_subsample: int = 1, cameras: Sequence[str] = NCLT_CAMERAS):
        self.sequence, self.root_dir, self.image_subsample = str(sequence), Path(root_dir), int(image_subsample)
        self.cameras = list(cameras)
        self.calib = self.read_calib()
        self._maps: Dict[str, dict] = {}
        self._params: Dict[str, dict] = {}
        self._times: Optional[dict] = None

    # ---- files ------------------------------------------------------------------------------------------------------
    def read_times(self) -> Dict[str, List[int]]:
        """Timestamps present both as a Cam1 image and as a synchronised scan (nclt.py:290-302)."""
        if self._times is None:
            img = {int(f.stem) for f in (self.root_dir / "images" / self.sequence / "lb3" / "Cam1").glob("*.tiff")}
            pcl = {int(f.stem) for f in (self.root_dir / "velodyne_data" / self.sequence / "velodyne_sync").glob("*.bin")}
            common = sorted(img & pcl)
            self._times = {"image": common, "pcl": common}
        return self._times

    @property
    def timestamps(self) -> List[float]:
        ts = self.read_times()["pcl"]
        return [(t - ts[0]) / 1e6 for t in ts]   # seconds since the first scan (nclt.py:304-312)

    def __len__(self) -> int:
        return len(self.read_times()["pcl"])

    def read_image_files(self, frame_id: int) -> Dict[str, Path]:
        ts = self.read_times()["image"][frame_id]
        return {c: self.root_dir / "images" / self.sequence / "lb3" / c / f"{ts}.tiff" for c in self.cameras}

    def pcl_file(self, frame_id: int) -> Path:
        return self.root_dir / "velodyne_data" / self.sequence / "velodyne_sync" / f"{self.read_times()['pcl'][frame_id]}.bin"

    # ---- point clouds -------------------------------------------------------------------------------------------------
    def read_pcl(self, frame_id: int = -1, filename=None) -> np.ndarray:
        """velodyne_sync record = 4 x int16 (x, y, z, intensity | laser id); metres = raw * 0.005 - 100; points at 50 m or
        more are dropped (nclt.py:114-151).  Returns (N, 3) float32."""
        if (frame_id == -1) == (filename is None):
            raise AssertionError("Either frame_id or filename must be provided")
        raw = np.fromfile(filename if filename is not None else self.pcl_file(frame_id), dtype=np.int16)
        raw = raw[: raw.size // 4 * 4].reshape(-1, 4)[:, :3]
        pcl = raw.astype(np.float32) * np.float32(0.005) + np.float32(-100.0)
        return pcl[np.linalg.norm(pcl, axis=1) < 50]

    # ---- calibration --------------------------------------------------------------------------------------------------
    def read_calib(self) -> Dict[str, np.ndarray]:
        lidar_in_ego = _pose([0.002, -0.004, -0.957], [0.807, 0.166, -90.703])   # nclt.py:156-161, from the dataset paper
        return {"lidar_in_ego": lidar_in_ego, "ego_in_lidar": np.linalg.inv(lidar_in_ego)}

    def read_camera_parameters(self, camera: str) -> Dict[str, np.ndarray]:
        """K_cam<N>.csv (3 x 3) and x_lb3_c<N>.csv (x, y, z, roll, pitch, yaw in degrees): nclt.py:203-219."""
        if camera not in self._params:
            n = camera[-1]
            K = np.loadtxt(self.root_dir / "cam_params" / f"K_cam{n}.csv", delimiter=",")
            x = np.loadtxt(self.root_dir / "cam_params" / f"x_lb3_c{n}.csv", delimiter=",")
            self._params[camera] = {"K": K, "x_lb3": _pose(x[:3], x[3:])}
        return self._params[camera]

    def read_undistortion_map(self, camera: str) -> dict:
        """U2D_<camera>_1616X1232.txt: a header with the frame size, then one line per pixel `row col v u`
        (nclt.py:164-185).  Returns the full maps and the maps of the cropped, 90-degree-rotated output image."""
        if camera in self._maps:
            return self._maps[camera]
        import cv2
        path = self.root_dir / "cam_params" / f"U2D_{camera}_1616X1232.txt"
        with open(path, "r") as f:
            header = f.readline()
            w, h = (int(c) for c in re.sub(r"[^0-9,]", "", header).split(",")[:2])
            body = np.fromstring(f.read(), sep=" ", dtype=np.float64)   # one vectorised pass over ~2 M `row col v u` lines
        body = body.reshape(-1, 4)
        mapu = np.zeros((h, w), dtype=np.float32)
        mapv = np.zeros((h, w), dtype=np.float32)
        r, c = body[:, 0].astype(np.int64), body[:, 1].astype(np.int64)
        mapu[r, c] = body[:, 3].astype(np.float32)
        mapv[r, c] = body[:, 2].astype(np.float32)
        y0, x0, ch, cw = NCLT_CROP
        # output pixel (i, j) of the clockwise-rotated crop is crop pixel (ch - 1 - j, i): rotate the maps instead of the image
        mu = np.ascontiguousarray(cv2.rotate(mapu[y0:y0 + ch, x0:x0 + cw], cv2.ROTATE_90_CLOCKWISE))
        mv = np.ascontiguousarray(cv2.rotate(mapv[y0:y0 + ch, x0:x0 + cw], cv2.ROTATE_90_CLOCKWISE))
        self._maps[camera] = {"mapu": mapu, "mapv": mapv, "mapu_out": mu, "mapv_out": mv}
        return self._maps[camera]

    # ---- images -------------------------------------------------------------------------------------------------------
    def read_images(self, frame_id: int = -1, crop: bool = True, filenames: Optional[Sequence] = None) -> Dict[str, np.ndarray]:
        """camera -> RGB uint8 image, undistorted, cropped to the LiDAR-covered window and rotated upright
        (nclt.py:68-112); (700 // s, 820 // s)... i.e. 700 rows x 820 columns at image_subsample 1."""
        import cv2
        if (frame_id == -1) == (filenames is None):
            raise AssertionError("Either frame_id or filenames must be provided")
        files = list(filenames) if filenames is not None else [self.read_image_files(frame_id)[c] for c in self.cameras]
        out = {}
        for camera, path in zip(self.cameras, files):
            bgr = cv2.imread(str(path))
            if bgr is None:
                raise FileNotFoundError(f"cannot read image {path}")
            m = self.read_undistortion_map(camera)
            if crop:
                img = cv2.remap(bgr, m["mapu_out"], m["mapv_out"], cv2.INTER_LINEAR)
            else:
                img = cv2.rotate(cv2.remap(bgr, m["mapu"], m["mapv"], cv2.INTER_LINEAR), cv2.ROTATE_90_CLOCKWISE)
            img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
            if self.image_subsample > 1:
                img = cv2.resize(img, (img.shape[1] // self.image_subsample, img.shape[0] // self.image_subsample),
                                 interpolation=cv2.INTER_AREA)
            out[camera] = img
        return out

    # ---- projection ---------------------------------------------------------------------------------------------------
    def camera_from_body(self, camera: str) -> np.ndarray:
        """T_c_body of nclt.py:317-327: inv(x_lb3_c) @ inv(x_body_lb3), the Ladybug pose from the dataset SDK."""
        x_body_lb3 = _pose([0.035, 0.002, -1.23], [-179.93, -0.23, 0.50])
        return np.linalg.inv(self.read_camera_parameters(camera)["x_lb3"]) @ np.linalg.inv(x_body_lb3)

    def camera_spec(self, camera: str, image_hw) -> api.CameraSpec:
        """What project_pcl_to_image (nclt.py:311-366) computes per call, as the parameters of the fused GPU kernel: projection
        K @ T_c_body, truncation before the window test, the crop window in sub-sampled pixels, black-pixel rejection.
        The stored image (and its token grid) is the ROTATED crop, (x, y) index the unrotated one (prepare_scenes.py:73-81):
        rot90, with img_hw describing the frame the projection lives in = the stored image's (width, height)."""
        s = self.image_subsample
        y0, x0, h, w = (v // s for v in NCLT_CROP)
        P = self.read_camera_parameters(camera)["K"] @ self.camera_from_body(camera)[:3]
        return api.CameraSpec(P=P, img_hw=(image_hw[1], image_hw[0]), grid_hw=(0, 0), crop=(y0, x0, h, w), subsample=float(s),
                              black_mode=1, rot90=True)

    def project_params(self, images: Dict[str, np.ndarray]) -> Dict[str, api.CameraSpec]:
        return {c: self.camera_spec(c, img.shape[:2]) for c, img in images.items()}


class OxfordRobotcar:
    """One Oxford Radar RobotCar sequence (reference: dataloader/oxford_robotcar.py).  ``sdk_dir`` is a robotcar-dataset-sdk
    checkout (extrinsics/*.txt, models/*.txt): the reference vendors it next to the loader, this package does not."""

    def __init__(self, sequence: str, root_dir, sdk_dir, image_subsample: int = 1, cameras: Sequence[str] = ROBOTCAR_CAMERAS):
        self.sequence, self.root_dir, self.sdk_dir = str(sequence), Path(root_dir), Path(sdk_dir)
        self.image_subsample, self.cameras = int(image_subsample), list(cameras)
        self.lidar_frequency = 10
        self._luts: Dict[str, np.ndarray] = {}
        self.calib = self.read_calib()

    @property
    def seq_dir(self) -> Path:
        return self.root_dir / f"{self.sequence}-radar-oxford-10k"

    def read_pcl(self, frame_id: int = -1, filename=None, timestamps: Optional[Sequence[int]] = None) -> np.ndarray:
        """velodyne_left/<ts>.bin = float32 (4, N): x, y, z, intensity; points within 2.5 m (the car) or at 50 m or more are
        dropped (oxford_robotcar.py:158-183).  Returns (N, 3) float32."""
        if (frame_id == -1) == (filename is None):
            raise AssertionError("Either frame_id or filename must be provided")
        if filename is None:
            if timestamps is None:
                raise ValueError("frame_id needs the list of scan timestamps")
            filename = self.seq_dir / "velodyne_left" / f"{timestamps[frame_id]}.bin"
        pcl = np.fromfile(filename, dtype=np.float32).reshape(4, -1).T
        depth = np.linalg.norm(pcl[:, :3], axis=1)
        return np.ascontiguousarray(pcl[(depth > 2.5) & (depth < 50), :3])

    def _extrinsics(self, name: str) -> np.ndarray:
        with open(self.sdk_dir / "extrinsics" / f"{name}.txt") as f:
            return se3_from_xyzrpy([float(x) for x in next(f).split(" ")])

    def read_calib(self) -> Dict[str, np.ndarray]:
        """oxford_robotcar.py:185-217: everything relative to the stereo/centre camera ("ego")."""
        calib = {"lidar_in_ego": self._extrinsics("velodyne_left")}
        for camera in self.cameras:
            calib[f"{camera}_in_ego"] = self._extrinsics("stereo" if camera == "stereo/centre" else camera)
        calib["ins_in_ego"] = self._extrinsics("ins")
        calib["lidar_in_ins"] = np.linalg.solve(calib["ins_in_ego"], calib["lidar_in_ego"])
        calib["ins_in_lidar"] = np.linalg.inv(calib["lidar_in_ins"])
        return calib

    def read_intrinsics(self, camera: str):
        """models/<camera>.txt of the SDK (camera_model.py:85-100): first line fx fy cx cy, then the 4 x 4 G_camera_image."""
        stem = {"stereo/centre": "stereo_narrow_left"}.get(camera, camera)
        with open(self.sdk_dir / "models" / f"{stem}.txt") as f:
            fx, fy, cx, cy = (float(x) for x in next(f).split())
            G = np.array([[float(x) for x in line.split()] for line in f if line.strip()], dtype=np.float64)
        return (fx, fy), (cx, cy), G

    def read_lut(self, camera: str) -> np.ndarray:
        """models/<model>_distortion_lut.bin (camera_model.py:148-154): float64 (2, H * W) = for every pixel of the undistorted
        image the (u, v) it comes from in the distorted one.  Returned as the (2, H*W) (row, column) coordinate array that
        ``undistort`` hands to the interpolation."""
        if camera not in self._luts:
            stem = {"stereo/centre": "stereo_narrow_left"}.get(camera, camera)
            lut = np.fromfile(self.sdk_dir / "models" / f"{stem}_distortion_lut.bin", np.double)
            lut = lut.reshape(2, lut.size // 2)
            self._luts[camera] = np.ascontiguousarray(lut[::-1])   # (v, u): row coordinates first
        return self._luts[camera]

    def undistort(self, camera: str, image: np.ndarray) -> np.ndarray:
        """CameraModel.undistort (camera_model.py:85-117): bilinear look-up of every channel through the LUT (the same
        scipy.ndimage.map_coordinates(order=1) call, so the same values), cast back to the image's dtype."""
        from scipy.ndimage import map_coordinates
        lut = self.read_lut(camera)
        if image.ndim != 3:
            raise ValueError("Undistortion function only works with multi-channel images")
        if image.shape[0] * image.shape[1] != lut.shape[1]:
            raise ValueError("Incorrect image size for camera model")
        coords = lut.reshape(2, image.shape[0], image.shape[1])
        return np.stack([map_coordinates(image[:, :, c], coords, order=1) for c in range(image.shape[2])], axis=2).astype(image.dtype)

    def read_images(self, filenames: Sequence, raw: bool = False) -> Dict[str, np.ndarray]:
        """camera -> RGB uint8 image.  ``raw=False``: the files are the undistorted, hood-cropped images the reference caches as
        <camera>_undistorted/<ts>.png.  ``raw=True``: the files are the Bayer PNGs of the dataset and go through the reference's
        own chain (oxford_robotcar.py:101-137): bilinear demosaic (GBRG for stereo/centre, RGGB for the mono cameras), LUT
        undistortion, cast to uint8, crop of the bonnet (150 rows) / of the area without LiDAR coverage (200 rows).  Then the
        optional bilinear sub-sampling."""
        from PIL import Image
        out = {}
        for camera, path in zip(self.cameras, filenames):
            img = Image.open(path)
            if raw:
                rgb = demosaic_bilinear(np.asarray(img, dtype=np.float64), "GBRG" if camera == "stereo/centre" else "RGGB")
                rgb = self.undistort(camera, rgb).astype(np.uint8)
                img = Image.fromarray(rgb)
                img = img.crop((0, 0, img.size[0], img.size[1] - (150 if camera == "stereo/centre" else 200)))
            if self.image_subsample > 1:
                img = img.resize((img.size[0] // self.image_subsample, img.size[1] // self.image_subsample), Image.BILINEAR)
            out[camera] = np.array(img)
        return out

    def camera_spec(self, camera: str, image_hw) -> api.CameraSpec:
        """project_pcl_to_image (oxford_robotcar.py:330-363): p' = solve(G_camera_image, cam_in_ego @ lidar_in_ego @ p), z >= 0,
        u = fx x / z + cx, v = fy y / z + cy, divided by the sub-sampling, bounds 0 <= u <= W and 0 <= v <= H tested on the
        float values (upper bound inclusive, as the reference has it), then truncation; a black pixel still claims the point,
        with a zero descriptor (prepare_scenes.py:57-62)."""
        (fx, fy), (cx, cy), G = self.read_intrinsics(camera)
        M = np.linalg.solve(G, self.calib[f"{camera}_in_ego"] @ self.calib["lidar_in_ego"])
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
        return api.CameraSpec(P=K @ M[:3], img_hw=tuple(image_hw), grid_hw=(0, 0), subsample=float(self.image_subsample),
                              z_inclusive=True, float_bounds=True, black_mode=2)

    def project_params(self, images: Dict[str, np.ndarray]) -> Dict[str, api.CameraSpec]:
        return {c: self.camera_spec(c, img.shape[:2]) for c, img in images.items()}
