"""Scene containers of the reference's experiment driver and the pre-processing it applies before the hot path
(SURVEY.md section 8f rows 2-3).  Host-side Python, like the reference; the voxel operations run on libvfmreg_b200.so.

  read_scene_json(file)             data/*/scene_*.json as written by the scene generator and read by prepare_scenes.py:122-131:
                                    {"mapping": {"point_clouds", "images", "poses"}, "registration": [{"point_cloud", "images", "pose"}]}
  read_scenes(file) / save_scene    the HDF5 layout of prepare_scenes.py:16-47 and vfm_reg/read_h5.py:17-49 (needs h5py, which the
                                    reference installs; it is not a dependency of the hot path and is imported lazily)
  build_local_map(poses, clouds)    registration_node.py:557-581: per map frame drop the points without descriptor, voxel-thin at
                                    0.25 m, move into the map frame; thin the concatenation again (in two halves above 10^6 points)
  prepare_scan(cloud)               registration_node.py:588-590: voxel-thin the scan at 0.1 m

Down-sampled rows come back in input order (the reference: the same rows in tsl::robin_map order)."""
from __future__ import annotations

import json
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Sequence

import numpy as np

from . import metrics
from .voxel import voxel_down_sample


@dataclass
class SceneSpec:
    """One scene_*.json: the frames that build the local map and the scans to be registered against it."""
    map_point_clouds: List[str]
    map_images: List[List[str]]
    map_poses: np.ndarray            # (F, 4, 4) float64
    scan_point_clouds: List[str]
    scan_images: List[List[str]]
    scan_poses: np.ndarray           # (S, 4, 4) float64


def read_scene_json(filename) -> SceneSpec:
    with open(filename, "r", encoding="utf-8") as f:
        d = json.load(f)
    m, reg = d["mapping"], d["registration"]
    if not (len(m["point_clouds"]) == len(m["images"]) == len(m["poses"])):
        raise ValueError("Invalid scene file: mapping lists differ in length")
    poses = np.asarray(m["poses"], dtype=np.float64).reshape(-1, 4, 4)
    scan_poses = np.asarray([r["pose"] for r in reg], dtype=np.float64).reshape(-1, 4, 4)
    return SceneSpec(list(m["point_clouds"]), [list(x) for x in m["images"]], poses, [r["point_cloud"] for r in reg],
                     [list(r["images"]) for r in reg], scan_poses)


def _h5py():
    try:
        import h5py
    except ImportError as e:  # pragma: no cover
        raise ImportError("the HDF5 scene files need h5py (the reference installs it; it is not needed by the hot path)") from e
    return h5py


def read_scenes(filename) -> Dict[str, list]:
    """vfm_reg/read_h5.py:17-49: {"map_poses", "map_point_clouds", "map_clip", "scene_poses", "scene_point_clouds"}."""
    h5py = _h5py()
    scene = {"map_poses": [], "map_point_clouds": [], "map_clip": [], "scene_poses": [], "scene_point_clouds": []}
    with h5py.File(filename, "r") as f:
        for key in f["map"].keys():   # one key: the mapping sequence
            g = f["map"][key]
            for pose, pc in zip(g["pose"].values(), g["point_cloud"].values()):
                scene["map_poses"].append(pose[()])
                scene["map_point_clouds"].append(pc[()])
            if "clip" in g.keys():
                scene["map_clip"].extend(c[()] for c in g["clip"].values())
        for scan in f["scans"]:
            scene["scene_poses"].append(f["scans"][scan]["pose"][()])
            scene["scene_point_clouds"].append(f["scans"][scan]["point_cloud"][()])
    return scene


def save_scene(filename, sequences: Sequence[str], map_poses, map_point_clouds, seq_poses, seq_point_clouds) -> None:
    """prepare_scenes.py:16-47."""
    h5py = _h5py()
    filename = Path(filename)
    filename.parent.mkdir(parents=True, exist_ok=True)
    with h5py.File(filename, "w") as f:
        g = f.create_group(f"map/{sequences[0]}")
        pg, cg = g.create_group("pose"), g.create_group("point_cloud")
        for j in range(len(map_poses)):
            pg.create_dataset(f"{j:03}", data=map_poses[j])
            cg.create_dataset(f"{j:03}", data=map_point_clouds[j])
        sg = f.create_group("scans")
        for j in range(len(seq_poses)):
            if seq_poses[j] is None:   # "This sequence has no hits"
                continue
            s = sg.create_group(f"{sequences[j + 1]}")
            s.create_dataset("pose", data=seq_poses[j])
            s.create_dataset("point_cloud", data=seq_point_clouds[j])


def build_local_map(map_poses, map_point_clouds, *, voxel_size: float = 0.25, feat_dim: int = 384, has_descriptor: str = "sum",
                    split_above: int = 1_000_000, device=None) -> np.ndarray:
    """The local map the reference registers against (registration_node.py:557-581), (M, 3 + feat_dim) float32.

    ``has_descriptor``: "sum" keeps the reference's test ``sum(descriptor) > 0`` (:562), "norm" keeps rows with a non-zero
    descriptor (with a zero-mean ChannelNorm output the sum test is decided by rounding noise)."""
    frames = []
    for pose, pcl in zip(map_poses, map_point_clouds):
        pcl = np.asarray(pcl)
        if has_descriptor == "sum":
            keep = np.sum(pcl[:, 3:], axis=1) > 0
        elif has_descriptor == "norm":
            keep = np.any(pcl[:, 3:] != 0, axis=1)
        else:
            raise ValueError(f"Invalid has_descriptor: {has_descriptor}")
        pcl = pcl[keep]
        if pcl.shape[0]:
            pcl = voxel_down_sample(pcl, voxel_size, device=device).astype(pcl.dtype)
        frames.append(metrics.transform_pcl(pcl, np.asarray(pose, dtype=np.float64)))
    local = np.concatenate(frames, axis=0).astype(np.float32)
    if local.shape[0] > split_above:   # "Split up voxelization due to memory constraints"
        mean_x = np.mean(local[:, :3], axis=0)[0]
        a = voxel_down_sample(local[local[:, 0] > mean_x], voxel_size, device=device)
        b = voxel_down_sample(local[local[:, 0] <= mean_x], voxel_size, device=device)
        local = np.concatenate([a, b], axis=0)
    elif local.shape[0]:
        local = voxel_down_sample(local, voxel_size, device=device)
    return local[:, :3 + feat_dim]


def prepare_scan(point_cloud, voxel_size: float = 0.1, device=None) -> np.ndarray:
    """registration_node.py:588-590."""
    point_cloud = np.asarray(point_cloud)
    return voxel_down_sample(point_cloud, voxel_size, device=device).astype(point_cloud.dtype)


def prepare_scene(dataset_dir, scene_file, output_dir, feature_generator, *, dataset: str = "", sdk_dir=None, image_subsample: int = 1,
                  device=None) -> Path:
    """The producer (prepare_scenes.py:110-167, `python prepare_scenes.py /data/DATASET /scenes/DATASET` for one scene file):
    every map frame is read, voxel-thinned at 0.2 m and given per-point descriptors from its surround images, every scan the
    same at 0.1 m, and the scene goes to ``output_dir/<scene>.h5``.  Differences in mechanics only: the voxel thinning and
    projection + gather run on the GPU (`voxel_down_sample`, `features.create_descriptors` -> one fused kernel per frame
    instead of the per-point Python loops of project_pcl_to_image / create_descriptors).

    ``dataset``: "nclt" or "robotcar" (default: from the directory name, as the reference decides); ``sdk_dir``: the RobotCar
    SDK checkout (extrinsics, camera models, LUTs); ``image_subsample`` stays at the loaders' default of 1 -- the reference
    sets a local `image_subsample = 2` (:119) but never passes it on."""
    from . import datasets, features
    dataset_dir, scene_file, output_dir = Path(dataset_dir), Path(scene_file), Path(output_dir)
    name = dataset or dataset_dir.name
    if "nclt" in name:
        make, date_idx = (lambda seq: datasets.NCLT(seq, dataset_dir, image_subsample)), 1
    elif "robotcar" in name:
        if sdk_dir is None:
            raise ValueError("Unknown dataset: RobotCar needs sdk_dir (extrinsics, camera models)")
        make, date_idx = (lambda seq: datasets.OxfordRobotcar(seq, dataset_dir, sdk_dir, image_subsample)), 0
    else:
        raise ValueError("Unknown dataset")   # prepare_scenes.py:117
    spec = read_scene_json(scene_file)
    sequences = [spec.map_point_clouds[date_idx].split("/")[1]] + [p.split("/")[date_idx] for p in spec.scan_point_clouds]   # :127-130
    loaders: Dict[str, object] = {}

    def frame(seq_name: str, pcl_file: str, image_files: Sequence[str], leaf: float) -> np.ndarray:
        seq = loaders.setdefault(seq_name, make(seq_name))   # one loader (calibration, undistortion maps) per sequence
        pcl = seq.read_pcl(filename=dataset_dir / pcl_file)
        pcl = voxel_down_sample(pcl, leaf, device=device).astype(pcl.dtype)
        files = [dataset_dir / f for f in image_files]
        images = seq.read_images(filenames=files) if isinstance(seq, datasets.NCLT) else seq.read_images(files, raw=True)
        desc = features.create_descriptors(images, seq.project_params(images), feature_generator, pcl)
        return np.c_[pcl, desc]

    map_clouds = [frame(sequences[0], p, imgs, 0.2) for p, imgs in zip(spec.map_point_clouds, spec.map_images)]       # :133-146
    scan_clouds = [frame(sequences[i + 1], p, imgs, 0.1) for i, (p, imgs) in enumerate(zip(spec.scan_point_clouds, spec.scan_images))]
    out = output_dir / scene_file.name.replace(".json", ".h5")
    save_scene(out, sequences, list(spec.map_poses), map_clouds, list(spec.scan_poses), scan_clouds)                   # :166-167
    return out
