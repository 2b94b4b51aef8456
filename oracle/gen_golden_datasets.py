#!/usr/bin/env python
"""Golden vectors for vfm_registration_b200/datasets.py: the reference's own NCLT loader (/root/reference/src/vfm-reg/src/
dataloader/nclt.py) executed on the miniature tree of tests/synth_dataset.py.  Test infrastructure (oracle/): run here, in the
container that has /root/reference; the outputs are committed as tests/golden/datasets_nclt.npz.
    python oracle/gen_golden_datasets.py"""
import hashlib
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, "/root/reference/src/vfm-reg/src")
import synth_dataset  # noqa: E402
from oracle import gen_golden  # noqa: E402

gen_golden._install_stubs()   # kiss_icp & co. (imported by the dataloader package, untouched by the NCLT class) become inert stubs

from dataloader.nclt import NCLT as RefNCLT  # noqa: E402

out = {}
with tempfile.TemporaryDirectory() as tmp:
    synth_dataset.build(tmp)
    for sub in (1, 2):
        ref = RefNCLT(synth_dataset.SEQ, Path(tmp), image_subsample=sub)
        imgs = ref.read_images(frame_id=0)
        pcl = ref.read_pcl(frame_id=0)
        if sub == 1:
            out["pcl"] = pcl
            out["lidar_in_ego"] = ref.calib["lidar_in_ego"]
            for cam in ("Cam1", "Cam2", "Cam5"):
                out[f"K_{cam}"] = ref.camera_parameters[cam]["K"]
                out[f"x_lb3_{cam}"] = ref.camera_parameters[cam]["x_lb3"]
            out["timestamps"] = np.asarray(ref.timestamps_abs["pcl"])
        for cam in ("Cam1", "Cam2"):
            im = imgs[cam]
            out[f"img{sub}_{cam}_shape"] = np.asarray(im.shape)
            out[f"img{sub}_{cam}_sha"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(im).tobytes()).digest(), dtype=np.uint8)
            out[f"img{sub}_{cam}_patch"] = im[100:116, 200:216].copy()
            # the projection the producer runs (prepare_scenes.py:66-76): homogeneous points, image rotated back
            import cv2
            p4 = np.insert(pcl.astype(np.float64), 3, values=1, axis=1).T
            x_im, y_im, idx = ref.project_pcl_to_image(p4, cv2.rotate(im, cv2.ROTATE_90_COUNTERCLOCKWISE), cam)
            out[f"proj{sub}_{cam}_x"], out[f"proj{sub}_{cam}_y"], out[f"proj{sub}_{cam}_idx"] = x_im, y_im, idx
# ---- RobotCar: the SDK's own CameraModel.undistort (robotcar_sdk/python/camera_model.py) on a demosaiced synthetic Bayer image.
# colour_demosaicing is not installable offline: the demosaic is the restatement in datasets.py (published bilinear kernels).
from dataloader.robotcar_sdk.python.camera_model import CameraModel  # noqa: E402
from vfm_registration_b200 import datasets as _ds  # noqa: E402
with tempfile.TemporaryDirectory() as tmp:
    h, w = 320, 416
    synth_dataset.write_robotcar_models(tmp, h, w)
    model = CameraModel(tmp, "/x/mono_left/")
    rgb = _ds.demosaic_bilinear(synth_dataset.robotcar_cfa(h, w).astype(np.float64), "RGGB")
    out["rc_hw"] = np.asarray([h, w])
    und = model.undistort(rgb)
    out["rc_undistorted_patch"] = und[40:56, 60:76].copy()
    out["rc_undistorted_sha"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(und).tobytes()).digest(), dtype=np.uint8)
os.makedirs(ROOT / "tests" / "golden", exist_ok=True)
np.savez_compressed(ROOT / "tests" / "golden" / "datasets_nclt.npz", **out)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
