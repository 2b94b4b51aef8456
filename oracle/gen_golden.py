#!/usr/bin/env python
"""Generates tests/golden/*.npz by EXECUTING the reference's own Python functions from
/root/reference (test infrastructure; runs only in the build container -- the GPU box has
no /root/reference and consumes the committed fixtures).

Third-party modules the reference imports but that are not installed offline (rospy, open3d,
faiss, h5py, kiss_icp, featup, ...) are replaced by inert stubs: none of them is touched by
the functions exercised here, which are pure NumPy / SciPy / torch / torchvision code:

  project_nclt.npz      NCLT.project_pcl_to_image              dataloader/nclt.py:311-366
  project_oxford.npz    OxfordRobotcar.project_pcl_to_image    dataloader/oxford_robotcar.py:330-363
  create_descriptors.npz create_descriptors                    prepare_scenes.py:50-107
  metrics.npz           compute_errors / compute_success_rate  registration_node.py:997-1025
                        transform_pcl                          vfm_reg/utils.py:47-54
  kabsch_pointdsc.npz   rigid_transform_3d                     pointdsc/common.py:7-47
  find_corr.npz         find_correspondences (nested)          registration_node.py:482-538
  preprocess.npz        create_transform_ + transform          vfm_reg/image_features.py:67-77,95
  upsample.npz          F.interpolate call of get_image_features  vfm_reg/image_features.py:104-108

Usage:  python oracle/gen_golden.py            (writes tests/golden/)
"""
from __future__ import annotations

import ast
import importlib.abc
import importlib.machinery
import os
import sys
import textwrap
import types
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference/src/vfm-reg/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


class _Stub(types.ModuleType):
    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Last-resort finder: any module nobody else can find becomes an inert stub."""

    STUBBED = {"rospy", "open3d", "faiss", "h5py", "kiss_icp", "featup", "hdbscan", "teaserpp_python", "sensor_msgs",
               "visualization_msgs", "geometry_msgs", "std_msgs", "tf_conversions", "pytorch_lightning", "matplotlib",
               "colour_demosaicing", "MinkowskiEngine", "gtsam", "pointnet2_ops"}

    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] not in self.STUBBED:
            return None
        return importlib.machinery.ModuleSpec(fullname, self, is_package=True)

    def create_module(self, spec):
        return _Stub(spec.name)

    def exec_module(self, module):
        pass


def _install_stubs():
    sys.meta_path.append(_StubFinder())
    for name in ("vfm_reg.descriptors", "pointdsc.PointDSC", "vfm_reg.read_h5", "dataloader.kitti_odometry"):
        sys.modules[name] = _Stub(name)
    sys.path.insert(0, REF)


def _nested_function(path: str, outer: str, name: str, glb: dict):
    """Extract a nested function's source from the reference file and compile it."""
    src = open(path).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == outer:
            for sub in ast.walk(node):
                if isinstance(sub, ast.FunctionDef) and sub.name == name:
                    code = textwrap.dedent(ast.get_source_segment(src, sub))
                    exec(compile(code, path, "exec"), glb)
                    return glb[name]
    raise RuntimeError(f"{name} not found in {path}")


def gen_project(rng):
    from dataloader.nclt import NCLT
    from dataloader.oxford_robotcar import OxfordRobotcar
    from scipy.spatial.transform import Rotation as R

    # ---- NCLT
    n = 4000
    pts = np.c_[rng.uniform(-20, 20, (n, 2)), rng.uniform(-2, 6, n)].astype(np.float32).astype(np.float64)
    pcl_h = np.insert(pts, 3, values=1, axis=1).T
    k = np.array([[205.0, 0.0, 160.3], [0.0, 203.5, 121.7], [0.0, 0.0, 1.0]])
    x_lb3 = np.eye(4)
    x_lb3[:3, :3] = R.from_euler("xyz", [88.0, 1.5, -92.0], degrees=True).as_matrix()
    x_lb3[:3, 3] = [0.02, -0.04, 0.1]
    coords = [20, 20, 200, 280]  # y0, x0, h, w (full resolution)
    sub = 2
    image = rng.integers(0, 255, (coords[2] // sub, coords[3] // sub, 3), dtype=np.uint8)
    image[10:25, 5:30] = 0
    image[40:50, 35:45, :2] = 0  # only two channels black -> still valid
    self = types.SimpleNamespace(cameras=["Cam1"], camera_parameters={"Cam1": {"K": k, "x_lb3": x_lb3}},
                                 image_subsample=sub, undistortion_masks={"Cam1": {"coords": coords}})
    x_im, y_im, idx = NCLT.project_pcl_to_image(self, pcl_h, image, "Cam1")
    # the fixed body->lb3 transform of nclt.py:320-323, as the reference builds it
    x_body_lb3 = np.eye(4)
    x_body_lb3[:3, 3] = [0.035, 0.002, -1.23]
    x_body_lb3[:3, :3] = R.from_euler("xyz", [-179.93, -0.23, 0.50], degrees=True).as_matrix()
    t_c_body = np.linalg.inv(x_lb3) @ np.linalg.inv(x_body_lb3)
    np.savez_compressed(os.path.join(OUT, "project_nclt.npz"), pts=pts, image=image, k=k, t_c_body=t_c_body,
                        coords=np.array(coords), sub=sub, x_im=x_im, y_im=y_im, idx=idx)
    print("project_nclt", len(idx), "of", n)

    # ---- Oxford
    n = 1500
    pts = np.c_[rng.uniform(-20, 20, (n, 2)), rng.uniform(-2, 6, n)].astype(np.float32).astype(np.float64)
    pcl_h = np.insert(pts, 3, values=1, axis=1).T
    lidar_in_ego = np.eye(4)
    lidar_in_ego[:3, :3] = R.from_euler("xyz", [0.5, -1.0, 2.0], degrees=True).as_matrix()
    lidar_in_ego[:3, 3] = [1.1, 0.0, -1.2]
    cam_in_ego = np.eye(4)
    cam_in_ego[:3, :3] = R.from_euler("xyz", [1.0, 0.3, -0.7], degrees=True).as_matrix()
    cam_in_ego[:3, 3] = [-1.7, 0.1, 1.0]
    g = np.array([[0.0, 0.0, 1.0, 0.0], [1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]])
    focal, principal, sub = (400.0, 401.5), (255.5, 191.25), 4
    image = rng.integers(0, 255, (96, 128, 3), dtype=np.uint8)
    cm = types.SimpleNamespace(G_camera_image=g, focal_length=focal, principal_point=principal)
    self = types.SimpleNamespace(cameras=["mono_left"], calib={"lidar_in_ego": lidar_in_ego, "mono_left_in_ego": cam_in_ego},
                                 camera_model={"mono_left": cm}, image_subsample=sub)
    u, v, idx = OxfordRobotcar.project_pcl_to_image(self, pcl_h, image, "mono_left")
    np.savez_compressed(os.path.join(OUT, "project_oxford.npz"), pts=pts, image=image, lidar_in_ego=lidar_in_ego,
                        cam_in_ego=cam_in_ego, g=g, focal=np.array(focal), principal=np.array(principal), sub=sub,
                        u=u, v=v, idx=idx)
    print("project_oxford", len(idx), "of", n)


def gen_create_descriptors(rng):
    import prepare_scenes
    from dataloader.nclt import NCLT

    n, c = 600, 8
    pts = np.c_[rng.uniform(-15, 15, (n, 2)), rng.uniform(-2, 4, n)].astype(np.float32)
    cams = ["c0", "c1", "c2"]
    hw = (40, 56)
    images = {cam: rng.integers(0, 255, (*hw, 3), dtype=np.uint8) for cam in cams}
    images["c0"][5:15, 10:30] = 0
    images["c2"][:, :8] = 0
    feats = {cam: rng.standard_normal((*hw, c)).astype(np.float32) for cam in cams}
    ks = {cam: np.array([[30.0 + 3 * i, 0, hw[1] / 2], [0, 31.0 - i, hw[0] / 2], [0, 0, 1.0]]) for i, cam in enumerate(cams)}
    from scipy.spatial.transform import Rotation as R
    ts = {}
    for i, cam in enumerate(cams):
        t = np.eye(4)
        # camera looks along +x rotated by yaw; overlapping fields of view so points are seen twice
        t[:3, :3] = (R.from_euler("z", 35.0 * i, degrees=True) * R.from_euler("yx", [90, -90], degrees=True)).as_matrix().T
        t[:3, 3] = [0.1 * i, 0.0, 0.2]
        ts[cam] = t

    def project(pcl_h, image, camera):
        q = ks[camera] @ (ts[camera] @ pcl_h)[:3]
        front = q[2] > 0
        x = (q[0] / q[2])[front].astype(int)
        y = (q[1] / q[2])[front].astype(int)
        inside = (x >= 0) & (x < image.shape[1]) & (y >= 0) & (y < image.shape[0])
        x, y = x[inside], y[inside]
        rgb = np.array([bool(np.any(image[y[i], x[i]] != 0)) for i in range(len(x))], dtype=bool)
        return x[rgb], y[rgb], np.where(front)[0][inside][rgb]

    class FakeGen:
        def get_image_features(self, image, upsample=True):
            for cam in cams:
                if image is images[cam]:
                    return feats[cam].copy()
            raise KeyError

    seq = types.SimpleNamespace(read_images=lambda filenames: images, project_pcl_to_image=project)
    out = prepare_scenes.create_descriptors(None, seq, FakeGen(), pts.copy())

    # NCLT branch (rot90 handling, prepare_scenes.py:73-74,80-81,93-94): square images so shapes survive
    hw2 = (48, 48)
    images2 = {cam: rng.integers(0, 255, (*hw2, 3), dtype=np.uint8) for cam in cams}
    images2["c1"][20:30, 0:20] = 0
    feats2 = {cam: rng.standard_normal((*hw2, c)).astype(np.float32) for cam in cams}

    class FakeGen2:
        def get_image_features(self, image, upsample=True):
            for cam in cams:
                if image is images2[cam]:
                    return feats2[cam].copy()
            raise KeyError

    seq2 = object.__new__(NCLT)
    seq2.read_images = lambda filenames: images2
    seq2.project_pcl_to_image = project
    out2 = prepare_scenes.create_descriptors(None, seq2, FakeGen2(), pts.copy())
    np.savez_compressed(os.path.join(OUT, "create_descriptors.npz"), pts=pts,
                        images=np.stack([images[c_] for c_ in cams]), feats=np.stack([feats[c_] for c_ in cams]),
                        ks=np.stack([ks[c_] for c_ in cams]), ts=np.stack([ts[c_] for c_ in cams]), out=out,
                        images2=np.stack([images2[c_] for c_ in cams]), feats2=np.stack([feats2[c_] for c_ in cams]),
                        out2=out2)
    print("create_descriptors nonzero rows", int((np.abs(out).sum(1) > 0).sum()), int((np.abs(out2).sum(1) > 0).sum()))


def gen_metrics(rng):
    import registration_node as rn
    from vfm_reg.utils import transform_pcl
    from scipy.spatial.transform import Rotation as R

    poses, gts, errs = [], [], []
    self = types.SimpleNamespace(rot_errors={}, trans_errors={})
    for i in range(32):
        p, g = np.eye(4), np.eye(4)
        p[:3, :3] = R.from_rotvec(rng.normal(0, 1.0, 3)).as_matrix()
        g[:3, :3] = p[:3, :3] @ R.from_rotvec(rng.normal(0, 0.05 * (i % 4), 3)).as_matrix()
        p[:3, 3] = rng.normal(0, 10, 3)
        g[:3, 3] = p[:3, 3] + rng.normal(0, 0.5 * (i % 3), 3)
        errs.append(rn.RegistrationNode.compute_errors(self, p, g, "vfm"))
        poses.append(p)
        gts.append(g)
    rates = [rn.RegistrationNode.compute_success_rate(self, "vfm", t, r) for t, r in ((0.3, 15), (0.6, 1.5), (2, 5), (1, 5))]
    pcl = rng.standard_normal((50, 3 + 5)).astype(np.float32)
    tp = transform_pcl(pcl, poses[3])
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), poses=np.stack(poses), gts=np.stack(gts), errs=np.array(errs),
                        rates=np.array(rates), pcl=pcl, pcl_t=tp)
    print("metrics rates", rates)


def gen_kabsch(rng):
    import torch
    from pointdsc.common import rigid_transform_3d

    a_list, b_list = [], []
    for i in range(24):
        k = 3 if i < 16 else 40
        a = rng.uniform(-30, 30, (k, 3))
        from scipy.spatial.transform import Rotation as R
        rot = R.from_rotvec(rng.normal(0, 1.5, 3)).as_matrix()
        b = a @ rot.T + rng.normal(0, 5, 3) + rng.normal(0, 0.05, (k, 3))
        if i % 4 == 3:  # mirrored target: forces the det(U)det(V) = -1 branch
            b = b * np.array([1.0, 1.0, -1.0])
        a_list.append(np.pad(a, ((0, 40 - k), (0, 0))))
        b_list.append(np.pad(b, ((0, 40 - k), (0, 0))))
    ts, ks = [], []
    for i, (a, b) in enumerate(zip(a_list, b_list)):
        k = 3 if i < 16 else 40
        t = rigid_transform_3d(torch.from_numpy(a[None, :k].astype(np.float32)), torch.from_numpy(b[None, :k].astype(np.float32)))
        ts.append(t[0].numpy())
        ks.append(k)
    np.savez_compressed(os.path.join(OUT, "kabsch_pointdsc.npz"), a=np.stack(a_list), b=np.stack(b_list), k=np.array(ks),
                        t=np.stack(ts))
    print("kabsch", len(ts))


def gen_find_corr(rng):
    from scipy.spatial import cKDTree
    fc = _nested_function(os.path.join(REF, "registration_node.py"), "compute_correspondences", "find_correspondences",
                          {"np": np, "cKDTree": cKDTree})
    f0 = rng.standard_normal((400, 32))
    f1 = np.concatenate([f0[rng.permutation(400)[:250]] + 0.1 * rng.standard_normal((250, 32)), rng.standard_normal((350, 32))])
    i0, i1 = fc(f0, f1, mutual_filter=True)
    j0, j1 = fc(f0, f1, n_points=100, mutual_filter=False)
    np.savez_compressed(os.path.join(OUT, "find_corr.npz"), f0=f0, f1=f1, i0=i0, i1=i1, j0=np.sort(j0), j1=j1[np.argsort(j0)])
    print("find_corr mutual", len(i0), "top-n", len(j0))


def gen_preprocess(rng):
    import torch
    import torch.nn.functional as F
    from vfm_reg.image_features import ImageFeatureGenerator

    self = types.SimpleNamespace(patch_size=14, patch_h=16, patch_w=None, transform=None, image_shape=[-1, -1])
    image = rng.integers(0, 255, (70, 82, 3), dtype=np.uint8)
    ImageFeatureGenerator.create_transform_(self, image.shape[0], image.shape[1])
    x = self.transform(image)
    sub = x[:, ::7, ::9].numpy().copy()
    np.savez_compressed(os.path.join(OUT, "preprocess.npz"), image=image, patch_w=self.patch_w, shape=np.array(x.shape),
                        sub=sub, total=float(x.double().sum()), abs_total=float(x.double().abs().sum()))
    print("preprocess", tuple(x.shape), self.patch_w)
    # the upsample of image_features.py:104-108 on a small token grid
    feat = torch.from_numpy(rng.standard_normal((1, 6, 16, 18)).astype(np.float32))
    up = F.interpolate(feat, (70, 82), mode="bilinear", align_corners=False)
    np.savez_compressed(os.path.join(OUT, "upsample.npz"), feat=feat[0].numpy(), up=up[0].numpy())


def main():
    os.makedirs(OUT, exist_ok=True)
    _install_stubs()
    rng = np.random.default_rng(20251017)
    gen_project(rng)
    gen_create_descriptors(rng)
    gen_metrics(rng)
    gen_kabsch(rng)
    gen_find_corr(rng)
    gen_preprocess(rng)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
