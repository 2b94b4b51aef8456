"""vfm_registration_b200 -- B200 (sm_100a) implementation of the descriptor-match-and-solve hot path of
vniclas/VFM-Registration behind a C ABI (include/vfmreg_b200.h).  See DESIGN.md."""
from .api import (CameraSpec, Context, MatchResult, RansacResult, RegResult, VfmRegError, filter_correspondences, get_context,  # noqa: F401
                  match_nn, project_gather, ransac_kabsch, register, register_batch, register_scans, ResidentMap, l2_distances, select_smallest, KdTree, ransac_nn_all, teaser_solve, TeaserResult)

__version__ = "0.1.0"
from .features import ImageFeatureGenerator, ViTFeaturizer, create_descriptors, extract_features  # noqa: E402,F401
from .voxel import VoxelMap, register_frame, register_frame_vfm, voxel_down_sample  # noqa: E402,F401
from . import datasets, scenes  # noqa: E402,F401
