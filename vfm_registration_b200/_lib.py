"""ctypes binding of libvfmreg_b200.so (include/vfmreg_b200.h).

The shared library is the product; this module only loads it and declares the argument types.  It fails loudly
when the library is missing or cannot be loaded -- there is no Python / CPU fallback for any of its entry points."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# VFMREG_LIB: load another build of the same library (kernel tuning experiments); the default is the in-tree build
LIB_PATH = os.environ.get("VFMREG_LIB") or os.path.join(HERE, "libvfmreg_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_NOGPU, ERR_ALLOC = 0, 1, 2, 3, 4
NORMALIZE, MUTUAL = 0x1, 0x2
ALGO_AUTO, ALGO_SIMT, ALGO_TC = 0x000, 0x100, 0x200


class VfmRegError(RuntimeError):
    pass


class RegisterParams(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("min_cos", C.c_float), ("ratio", C.c_float), ("n_hyp", C.c_int32),
                ("refit", C.c_int32), ("inlier_thresh", C.c_double), ("seed", C.c_uint64)]


class RegisterResult(C.Structure):
    _fields_ = [("T", C.c_double * 16), ("best_hyp", C.c_int64), ("n_inliers", C.c_int64), ("sumq", C.c_int64),
                ("n_corr", C.c_int64), ("fitness", C.c_double), ("rmse", C.c_double)]


class Camera(C.Structure):
    _fields_ = [("P", C.c_double * 12), ("img_h", C.c_int32), ("img_w", C.c_int32), ("crop_y0", C.c_int32),
                ("crop_x0", C.c_int32), ("crop_h", C.c_int32), ("crop_w", C.c_int32), ("grid_h", C.c_int32),
                ("grid_w", C.c_int32), ("subsample", C.c_double), ("z_inclusive", C.c_int32),
                ("float_bounds", C.c_int32), ("black_mode", C.c_int32), ("rot90", C.c_int32)]


class VitConfig(C.Structure):
    _fields_ = [("depth", C.c_int32), ("width", C.c_int32), ("heads", C.c_int32), ("mlp_dim", C.c_int32), ("patch", C.c_int32),
                ("patch_h", C.c_int32), ("channel_norm", C.c_int32), ("ln_eps", C.c_float), ("cn_eps", C.c_float),
                ("mean", C.c_float * 3), ("std", C.c_float * 3)]


class TeaserParams(C.Structure):
    _fields_ = [("noise_bound", C.c_double), ("cbar2", C.c_double), ("gnc_factor", C.c_double), ("cost_threshold", C.c_double),
                ("max_iterations", C.c_int32), ("reserved", C.c_int32), ("max_clique_nodes", C.c_int64)]


_P = C.c_void_p
_SIGS = {
    "vfmreg_teaser_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(TeaserParams), C.c_void_p, C.c_void_p, C.c_void_p]),
    "vfmreg_version": (C.c_int, []),
    "vfmreg_last_error": (C.c_char_p, []),
    "vfmreg_device_count": (C.c_int, []),
    "vfmreg_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "vfmreg_set_lanes": (C.c_int, [_P, C.c_int]),
    "vfmreg_destroy": (None, [_P]),
    "vfmreg_set_stream": (C.c_int, [_P, _P]),
    "vfmreg_sync": (C.c_int, [_P]),
    "vfmreg_kernel_launches": (C.c_int64, [_P]),
    "vfmreg_enable_timing": (C.c_int, [_P, C.c_int]),
    "vfmreg_group_time_ms": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "vfmreg_match_nn": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int64, C.c_int32, C.c_uint32, _P, _P, _P, _P, _P, _P]),
    "vfmreg_filter_correspondences": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_int, _P, _P]),
    "vfmreg_l2_distances": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "vfmreg_select_smallest": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64, _P, _P, _P]),
    "vfmreg_ransac": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, C.c_int32, _P, C.c_int32, C.c_uint64, C.c_double, C.c_int,
                                _P, _P, _P, _P, _P]),
    "vfmreg_register": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int32, C.POINTER(RegisterParams), _P, _P,
                                  _P, C.POINTER(RegisterResult)]),
    "vfmreg_register_host": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int32, C.POINTER(RegisterParams), _P,
                                       _P, _P, C.POINTER(RegisterResult)]),
    "vfmreg_register_batch": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, _P, _P, C.c_int32, C.POINTER(RegisterParams), _P, _P, _P, _P]),
    "vfmreg_register_batch_host": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, _P, _P, C.c_int32, C.POINTER(RegisterParams), _P, _P, _P,
                                             _P]),
    "vfmreg_map_create": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_uint32, C.c_int32, C.POINTER(_P)]),
    "vfmreg_map_destroy": (None, [_P]),
    "vfmreg_map_size": (C.c_int64, [_P]),
    "vfmreg_map_match": (C.c_int, [_P, _P, _P, C.c_int64, C.c_float, _P, _P, _P]),
    "vfmreg_register_scans": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, C.POINTER(RegisterParams), _P, C.c_int32, _P, _P, _P]),
    "vfmreg_project_gather": (C.c_int, [_P, _P, C.c_int64, C.POINTER(Camera), C.c_int32, _P, C.POINTER(C.c_int64), _P,
                                        C.POINTER(C.c_int64), C.c_int32, _P, _P, _P]),
    "vfmreg_vit_create": (C.c_int, [_P, C.POINTER(VitConfig), C.POINTER(_P)]),
    "vfmreg_vit_destroy": (None, [_P]),
    "vfmreg_vit_set_weight": (C.c_int, [_P, C.c_char_p, _P, C.c_int64]),
    "vfmreg_vit_set_pos_embed": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "vfmreg_vit_grid": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "vfmreg_vit_set_graphs": (C.c_int, [_P, C.c_int32]),
    "vfmreg_vit_forward": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "vfmreg_voxel_downsample": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_double, _P, _P]),
    "vfmreg_gather_rows": (C.c_int, [_P, _P, C.c_int32, _P, _P, C.c_int64, _P]),
    "vfmreg_voxel_map_create": (C.c_int, [_P, C.c_double, C.c_int32, C.POINTER(_P)]),
    "vfmreg_voxel_map_destroy": (None, [_P]),
    "vfmreg_voxel_map_build": (C.c_int, [_P, _P, _P, C.c_int64]),
    "vfmreg_voxel_map_size": (C.c_int64, [_P]),
    "vfmreg_voxel_map_points": (C.c_int, [_P, _P, _P, _P]),
    "vfmreg_voxel_map_nearest": (C.c_int, [_P, _P, _P, C.c_int64, C.c_double, _P, _P]),
    "vfmreg_kdtree_create": (C.c_int, [_P, _P, C.c_int64, C.POINTER(_P)]),
    "vfmreg_kdtree_destroy": (None, [_P]),
    "vfmreg_kdtree_nearest": (C.c_int, [_P, _P, _P, C.c_int64, C.c_double, _P, _P]),
    "vfmreg_ransac_nn_all": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, C.c_int, _P, _P, C.c_int32, _P, C.c_int32, C.c_uint64, C.c_double,
                                       _P, _P, _P, _P]),
    "vfmreg_register_frame": (C.c_int, [_P, _P, _P, C.c_int64, _P, C.c_double, C.c_double, C.c_int32, _P,
                                        C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "vfmreg_register_frame_vfm": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, C.c_int64, _P, C.c_double, C.c_double, C.c_int32, _P,
                                            C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
}

_lib = None


def declared_symbols():
    return sorted(_SIGS)


def load():
    """Load libvfmreg_b200.so; raises VfmRegError (never falls back) when it is absent or broken."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VfmRegError(f"{LIB_PATH} is missing: build it with `python -m vfm_registration_b200.build` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise VfmRegError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in _SIGS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise VfmRegError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != OK:
        msg = load().vfmreg_last_error().decode(errors="replace")
        raise VfmRegError(f"{what or 'libvfmreg_b200'} failed (code {rc}): {msg}")
