"""scenes.read_scenes / save_scene (prepare_scenes.py:16-47, vfm_reg/read_h5.py:17-49).  h5py is not installed in this image:
the round trip is executed against a minimal in-memory stand-in of the h5py calls the two functions make (groups as ordered
dicts, datasets answering `[()]`), and against the real library whenever it is importable."""
import sys
import types

import numpy as np
import pytest


class _Dataset:
    def __init__(self, data):
        self._a = np.array(data)

    def __getitem__(self, key):
        assert key == ()
        return self._a


class _Group(dict):
    def create_group(self, path):
        g = self
        for part in path.split("/"):
            g = g.setdefault(part, _Group())
        return g

    def create_dataset(self, name, data=None):
        self[name] = _Dataset(data)
        return self[name]

    def __getitem__(self, path):
        g = self
        for part in path.split("/"):
            g = dict.__getitem__(g, part)
        return g


_STORE = {}


class _File(_Group):
    def __init__(self, filename, mode):
        super().__init__()
        self._name, self._mode = str(filename), mode
        if mode == "r":
            self.update(_STORE[self._name])

    def __enter__(self):
        return self

    def __exit__(self, *a):
        if self._mode == "w":
            _STORE[self._name] = _Group(self)
        return False


def _roundtrip(tmp_path):
    from vfm_registration_b200 import scenes
    rng = np.random.default_rng(0)
    seqs = ["2012-01-08", "2012-02-04", "2012-03-17", "2012-05-26"]
    map_poses = [np.eye(4) + 0.01 * k for k in range(3)]
    map_pcs = [rng.standard_normal((50 + k, 387)).astype(np.float32) for k in range(3)]
    scan_poses = [np.eye(4) * 2, None, np.eye(4) * 3]            # the second sequence "has no hits" (prepare_scenes.py:37-38)
    scan_pcs = [rng.standard_normal((40, 387)).astype(np.float32), None, rng.standard_normal((41, 387)).astype(np.float32)]
    f = tmp_path / "processed" / "scene_000.h5"
    scenes.save_scene(f, seqs, map_poses, map_pcs, scan_poses, scan_pcs)
    s = scenes.read_scenes(f)
    assert len(s["map_poses"]) == 3 and len(s["scene_poses"]) == 2 and s["map_clip"] == []
    for a, b in zip(s["map_poses"], map_poses):
        assert np.array_equal(a, b)
    for a, b in zip(s["map_point_clouds"], map_pcs):
        assert np.array_equal(a, b)
    assert np.array_equal(s["scene_poses"][0], scan_poses[0]) and np.array_equal(s["scene_poses"][1], scan_poses[2])
    assert np.array_equal(s["scene_point_clouds"][1], scan_pcs[2])


def test_scene_h5_roundtrip_against_stand_in(tmp_path, monkeypatch):
    fake = types.ModuleType("h5py")
    fake.File = _File
    monkeypatch.setitem(sys.modules, "h5py", fake)
    _roundtrip(tmp_path)


def test_scene_h5_roundtrip_real_h5py(tmp_path):
    pytest.importorskip("h5py")
    _roundtrip(tmp_path)
