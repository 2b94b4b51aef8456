/*
 * vfmreg_b200.h -- C ABI of libvfmreg_b200.so: the B200 (sm_100a) implementation of the
 * descriptor-match-and-solve hot path of vniclas/VFM-Registration.
 *
 * Plain pointers and sizes only (no torch / pybind types).  Every entry point names the
 * reference interface it replaces (paths relative to the reference checkout); the
 * reference-side bindings a maintainer would add are shown in INTEGRATION.md.
 *
 * Conventions
 *   - All functions return VFMREG_OK (0) or an error code; vfmreg_last_error() gives the text
 *     (thread-local).  Nothing here falls back to a CPU path: without a CUDA device
 *     vfmreg_create() fails with VFMREG_ERR_NOGPU.
 *   - "device pointer" arguments are caller-owned, contiguous, row-major CUDA allocations on
 *     the context's device; the library never frees or retains them past the call's stream
 *     ordering.  Work is enqueued on the context's stream (vfmreg_set_stream) and is
 *     asynchronous unless stated otherwise.
 *   - Indices are int32 in CALLER ARRAY ORDER (the reference's indices are relative to
 *     tsl::robin_map iteration order and never leave C++, VoxelHashMap.cpp:662-676).
 */
#ifndef VFMREG_B200_H
#define VFMREG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VFMREG_VERSION 200

#if defined(__GNUC__)
#define VFMREG_API __attribute__((visibility("default")))
#else
#define VFMREG_API
#endif

enum {
  VFMREG_OK = 0,
  VFMREG_ERR_ARG = 1,    /* bad shape / null pointer / unsupported size (reference: ValueError("Invalid shape"), mapping.py:72-73) */
  VFMREG_ERR_CUDA = 2,   /* a CUDA call or kernel launch failed */
  VFMREG_ERR_NOGPU = 3,  /* no usable sm_100 device: there is no CPU fallback */
  VFMREG_ERR_ALLOC = 4
};

/* flags for vfmreg_match_nn / vfmreg_register */
#define VFMREG_NORMALIZE   0x1u   /* fvec_renorm_L2 both sides first (VoxelHashMap.cpp:474,480) */
#define VFMREG_MUTUAL      0x2u   /* also search b->a; register() keeps mutual pairs only (registration_node.py:530) */
#define VFMREG_ALGO_MASK   0xF00u
#define VFMREG_ALGO_AUTO   0x000u
#define VFMREG_ALGO_SIMT   0x100u /* exact fp32 CUDA-core kernel */
#define VFMREG_ALGO_TC     0x200u /* tcgen05 fp16 candidate search + exact fp32 re-rank (bit-identical results) */

typedef struct vfmreg_ctx vfmreg_ctx;

VFMREG_API int vfmreg_version(void);
VFMREG_API const char* vfmreg_last_error(void);
VFMREG_API int vfmreg_device_count(void);

/* One context per (process, device): owns a stream-ordered scratch arena. */
VFMREG_API int vfmreg_create(int device, vfmreg_ctx** ctx);
VFMREG_API void vfmreg_destroy(vfmreg_ctx* ctx);
VFMREG_API int vfmreg_set_stream(vfmreg_ctx* ctx, void* cuda_stream); /* cudaStream_t; NULL = legacy default stream */
VFMREG_API int vfmreg_sync(vfmreg_ctx* ctx);
/* vfmreg_register_batch spreads consecutive pairs over `lanes` CUDA streams (lane 0 = the context's stream, the others
 * are internal; fork/join by events, results are ordered as given).  The match kernel of one pair then runs beside the
 * latency-bound small kernels (filters, re-rank, RANSAC) of its neighbours.  1 <= lanes <= 8, default 5 (measured on configs[1]: 1 lane 1802 pairs/s, 3 lanes 2341, 5 lanes 2425, 8 lanes 2435). */
VFMREG_API int vfmreg_set_lanes(vfmreg_ctx* ctx, int lanes);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
VFMREG_API int64_t vfmreg_kernel_launches(const vfmreg_ctx* ctx);
/* device time in ms of the most recent launch of the named kernel group, measured with CUDA events on the
 * context's stream when profiling is enabled: group 0 = match GEMM (full searches), 1 = ransac score, 2 = project/gather, 3 = vit gemm,
 * 4 = match GEMM launches of the pruned reverse search of the mutual check (row count on the device) */
VFMREG_API int vfmreg_enable_timing(vfmreg_ctx* ctx, int on);
VFMREG_API int vfmreg_group_time_ms(vfmreg_ctx* ctx, int group, float* ms_total, int* launches);

/* ---------------------------------------------------------------------------------------------
 * a6/a7/a8  descriptor top-1 search.
 * Replaces VoxelHashMap::GetVFMCorrespondences' faiss block (VoxelHashMap.cpp:469-496: float32
 * renormalise, IndexFlatIP::add, search(k=1)) and, with VFMREG_MUTUAL, the two cKDTree queries of
 * find_correspondences (registration_node.py:487-527).
 *   a (n x d), b (m x d) float32 device; outputs device arrays:
 *   idx01[n], sim01[n] = argmax_j <a_i,b_j> (lowest j on ties) and its value,
 *   sec01[n] = runner-up value (may be NULL); idx10/sim10/sec10 [m] likewise for b->a (MUTUAL only, may be NULL).
 * Values are the canonical fp32 inner products (DESIGN.md "Canonical arithmetic").
 * ------------------------------------------------------------------------------------------- */
VFMREG_API int vfmreg_match_nn(vfmreg_ctx* ctx, const float* a, int64_t n, const float* b, int64_t m, int32_t d, uint32_t flags,
                    int32_t* idx01, float* sim01, float* sec01, int32_t* idx10, float* sim10, float* sec10);

/* Cosine gate (keep sim >= min_cos, VoxelHashMap.cpp:501-511), mutual check (registration_node.py:530) and
 * Lowe ratio test ((1-s1) < ratio^2 (1-s2)); pass NAN to disable min_cos / ratio.  Emits the kept
 * (query, match) pairs in query order into corr[n][2] and their number into *count (both device). */
VFMREG_API int vfmreg_filter_correspondences(vfmreg_ctx* ctx, const int32_t* idx01, const float* sim01, const float* sec01,
                                  const int32_t* idx10, int64_t n, float min_cos, float ratio, int mutual,
                                  int32_t* corr, int32_t* count);

/* a8 / a7 remainder.  The reference's brute-force block turns inner products of unit descriptors into L2 distances,
 * sqrt(2 - 2 a.b + 1e-6), takes the per-row argmin (registration_node.py:191-209) and then keeps the n_points smallest
 * distances (np.argpartition, :212-214; the non-mutual branch of find_correspondences does the same, :510-518).
 *   vfmreg_l2_distances   dist[i] = sqrt(2 - 2 sim01[i] + 1e-6) (float32; +inf where idx01[i] < 0; idx01 may be NULL)
 *   vfmreg_select_smallest  keeps the n_keep queries with the smallest distance (= largest similarity; ties at the
 *                         boundary towards the lowest query index -- the reference leaves them to introselect) and emits
 *                         their (query, match) pairs in query order into corr[n][2], their distances into dist (optional,
 *                         NULL to skip) and their number into *count.  All pointers device. */
VFMREG_API int vfmreg_l2_distances(vfmreg_ctx* ctx, const int32_t* idx01, const float* sim01, int64_t n, float* dist);
VFMREG_API int vfmreg_select_smallest(vfmreg_ctx* ctx, const int32_t* idx01, const float* sim01, int64_t n, int64_t n_keep,
                                      int32_t* corr, float* dist, int32_t* count);

/* ---------------------------------------------------------------------------------------------
 * a10  RANSAC on a correspondence list.
 * Replaces o3d.pipelines.registration.registration_ransac_based_on_correspondence(src, tgt, corres, max_dist,
 * TransformationEstimationPointToPoint(False), ransac_n=3, RANSACConvergenceCriteria(n_hyp, 1))
 * (call site registration_node.py:319-327).
 *   src_xyz (n x 3), tgt_xyz (m x 3): device, float32 (xyz_f64 = 0) or float64 (1)
 *   corr (max_corr x 2) int32 device; count: device int32 holding the number of valid rows (<= max_corr)
 *   sample_idx: device (n_hyp x 3) int32 indices into corr, or NULL -> drawn on the device from `seed`
 *   thresh: inlier distance tau (the reference passes 10000); refit != 0: least-squares refit on the inliers
 * Outputs (device; any of counts/sumq/mask may be NULL):
 *   T[16] row-major 4x4 float64; counts[n_hyp] inlier count per hypothesis (-1 = degenerate sample);
 *   sumq[n_hyp] = sum of rint(d^2 * 2^40 / tau^2) over inliers; mask[max_corr] inlier mask of the winner;
 *   stats[4] int64 = {best hypothesis or -1, its inlier count, its sumq, number of correspondences}.
 * ------------------------------------------------------------------------------------------- */
VFMREG_API int vfmreg_ransac(vfmreg_ctx* ctx, const void* src_xyz, const void* tgt_xyz, int xyz_f64, const int32_t* corr,
                  const int32_t* count, int32_t max_corr, const int32_t* sample_idx, int32_t n_hyp, uint64_t seed,
                  double thresh, int refit, double* T, int32_t* counts, int64_t* sumq, uint8_t* mask, int64_t* stats);

/* ---------------------------------------------------------------------------------------------
 * The whole consumer path = RegistrationNode.ransac_registration(method='vfm') without ICP
 * (registration_node.py:273-328): match -> gate -> RANSAC.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t flags;        /* VFMREG_NORMALIZE | VFMREG_MUTUAL | VFMREG_ALGO_* */
  float min_cos;         /* NAN = no gate; the reference uses 0.8 (registration_node.py:418) */
  float ratio;           /* NAN = no ratio test */
  int32_t n_hyp;         /* the reference uses 50000 (registration_node.py:326) */
  int32_t refit;
  double inlier_thresh;  /* the reference uses 10000 (registration_node.py:323) */
  uint64_t seed;
} vfmreg_register_params;

typedef struct {
  double T[16];          /* row-major 4x4, maps source into target */
  int64_t best_hyp;
  int64_t n_inliers;
  int64_t sumq;
  int64_t n_corr;
  double fitness;        /* n_inliers / n_corr */
  double rmse;           /* over the inliers of the returned hypothesis */
} vfmreg_register_result;

/* Device-resident inputs (src_xyz n x 3, tgt_xyz m x 3, src_feats n x d, tgt_feats m x d, all float32).
 * sample_idx: device (n_hyp x 3) or NULL.  corr_out (n x 2 int32) / mask_out (n uint8): device, optional.
 * Blocks until the result struct (host memory) is filled. */
VFMREG_API int vfmreg_register(vfmreg_ctx* ctx, const float* src_xyz, const float* tgt_xyz, const float* src_feats,
                    const float* tgt_feats, int64_t n, int64_t m, int32_t d, const vfmreg_register_params* params,
                    const int32_t* sample_idx, int32_t* corr_out, uint8_t* mask_out, vfmreg_register_result* result);

/* Same with HOST buffers: inputs are copied host->device and corr/mask device->host inside the call
 * (the end-to-end path a NumPy caller such as the reference's registration_node.py exercises). */
VFMREG_API int vfmreg_register_host(vfmreg_ctx* ctx, const float* src_xyz, const float* tgt_xyz, const float* src_feats,
                         const float* tgt_feats, int64_t n, int64_t m, int32_t d, const vfmreg_register_params* params,
                         const int32_t* sample_idx, int32_t* corr_out, uint8_t* mask_out,
                         vfmreg_register_result* result);

/* A batch of independent pairs with DEVICE-resident inputs: all pairs are enqueued back to back on the context's stream
 * and the host synchronises once at the end (no per-pair round trip).  corr_out[i] (n[i] x 2) / mask_out[i] (n[i]) are
 * device pointers or NULL. */
VFMREG_API int vfmreg_register_batch(vfmreg_ctx* ctx, int32_t n_pairs, const float* const* src_xyz, const float* const* tgt_xyz,
                          const float* const* src_feats, const float* const* tgt_feats, const int64_t* n, const int64_t* m,
                          int32_t d, const vfmreg_register_params* params, const int32_t* const* sample_idx,
                          int32_t* const* corr_out, uint8_t* const* mask_out, vfmreg_register_result* results);

/* A batch of independent pairs with HOST buffers (what registration_node.py's scene loop, :587-588, iterates serially):
 * the host->device copy of pair i+1 runs on a second stream while pair i is being matched and solved (double-buffered
 * staging), results come back in one transfer at the end.  Arrays of n_pairs pointers / sizes; sample_idx, corr_out,
 * mask_out (and their elements) may be NULL.  Pinned host memory makes the copies truly asynchronous. */
VFMREG_API int vfmreg_register_batch_host(vfmreg_ctx* ctx, int32_t n_pairs, const float* const* src_xyz, const float* const* tgt_xyz,
                               const float* const* src_feats, const float* const* tgt_feats, const int64_t* n, const int64_t* m,
                               int32_t d, const vfmreg_register_params* params, const int32_t* const* sample_idx,
                               int32_t* const* corr_out, uint8_t* const* mask_out, vfmreg_register_result* results);

/* Both batch entry points prepare (upload, renormalise, convert to fp16) a map ONCE for every run of consecutive pairs that
 * pass the same tgt_xyz / tgt_feats pointers and m: the reference builds one local map per scene and registers the scene's
 * 3-5 scans against it (registration_node.py:554-590). */

/* ---------------------------------------------------------------------------------------------
 * A map kept resident on the device across calls: what registration_node.py:554-580 builds once per scene
 * (`local_map`) and VoxelHashMap keeps in `map_n_` (VoxelHashMap.cpp:746-757) for the scans that follow.
 *   vfmreg_map_create    copies tgt_xyz (m x 3) and prepares tgt_feats (m x d) -- float32, HOST (host_buffers != 0) or
 *                        device pointers; flags = VFMREG_NORMALIZE | VFMREG_ALGO_*; returns after the map is ready.
 *   vfmreg_register_scans = vfmreg_register_batch[_host] with every pair's target = the resident map (params->flags must
 *                        carry the NORMALIZE / ALGO bits the map was created with).
 *   vfmreg_map_match     GetVFMCorrespondences' search against the resident map (VoxelHashMap.cpp:486-495): device queries
 *                        (n x d), outputs as vfmreg_match_nn's idx01 / sim01 / sec01.  min_cos (NAN = none) may be passed
 *                        when sec01 is NULL and the caller drops matches below it (VoxelHashMap.cpp:501-511): queries whose
 *                        best match is below the gate may then report idx01 = -1, sim01 = -inf instead of that match.
 * ------------------------------------------------------------------------------------------- */
typedef struct vfmreg_map vfmreg_map;
VFMREG_API int vfmreg_map_create(vfmreg_ctx* ctx, const float* tgt_xyz, const float* tgt_feats, int64_t m, int32_t d, uint32_t flags,
                                 int32_t host_buffers, vfmreg_map** map);
VFMREG_API void vfmreg_map_destroy(vfmreg_map* map);
VFMREG_API int64_t vfmreg_map_size(const vfmreg_map* map);
VFMREG_API int vfmreg_map_match(vfmreg_ctx* ctx, const vfmreg_map* map, const float* queries, int64_t n, float min_cos,
                                int32_t* idx01, float* sim01, float* sec01);
VFMREG_API int vfmreg_register_scans(vfmreg_ctx* ctx, const vfmreg_map* map, int32_t n_scans, const float* const* src_xyz,
                                     const float* const* src_feats, const int64_t* n, const vfmreg_register_params* params,
                                     const int32_t* const* sample_idx, int32_t host_buffers, int32_t* const* corr_out,
                                     uint8_t* const* mask_out, vfmreg_register_result* results);

/* ---------------------------------------------------------------------------------------------
 * a3/a4/a5  point -> pixel projection + feature gather + first-camera-wins scatter.
 * Replaces project_pcl_to_image (dataloader/nclt.py:311-366, dataloader/oxford_robotcar.py:330-363) and
 * create_descriptors' gather/dedup/scatter (prepare_scenes.py:57-104) in one pass, sampling the ViT token grid
 * directly instead of materialising the full-resolution map of image_features.py:104-108.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  double P[12];          /* 3x4 row-major: pixel_h = P * [x y z 1]^T in FULL-resolution pixels (K * T_cam_from_lidar) */
  int32_t img_h, img_w;  /* image size (after subsampling) the black-pixel test and the feature map refer to */
  int32_t crop_y0, crop_x0, crop_h, crop_w; /* window in subsampled pixels; image/feature origin = window origin */
  int32_t grid_h, grid_w;/* token grid of this camera */
  double subsample;      /* pixel = P-projection / subsample, then truncated toward zero (nclt.py:334-343) */
  int32_t z_inclusive;   /* 0: keep z > 0 (nclt.py:337); 1: keep z >= 0 (oxford_robotcar.py:344) */
  int32_t float_bounds;  /* 1: Oxford order -- test 0 <= u <= W, 0 <= v <= H on the un-truncated coordinate, then truncate
                            (oxford_robotcar.py:356-362; u == W / v == H would index out of range in the reference and is
                            dropped here); 0: NCLT order -- truncate, then test the crop window (nclt.py:342-347) */
  int32_t black_mode;    /* 0: ignore the image; 1: a black pixel hides the point from this camera (nclt.py:354-359);
                            2: a black pixel claims the point with a zero descriptor (prepare_scenes.py:58-62 on Oxford) */
  int32_t rot90;         /* 1: (u, v) live in the np.rot90(k=1) frame of the stored image / token grid
                            (prepare_scenes.py:73-74,80-81); img_h/img_w/crop then describe the rotated frame */
} vfmreg_camera;

/* points (n x 3) f32 device; cams: HOST array of n_cam descriptors; tokens: device f32, camera c's grid at
 * tokens + token_offsets[c] (HOST int64 array, in floats), layout (grid_h, grid_w, d);
 * images: device uint8 HWC RGB per camera at images + image_offsets[c] (HOST int64, bytes), or NULL to skip
 * the black-pixel test.  mode 0: bilinear sample of the token grid == nearest pixel of the align_corners=False
 * upsampled map (image_features.py:104-108 then prepare_scenes.py:85-91).
 * Outputs (device): desc (n x d) f32, zeros for unseen points; cam_of_point[n] int32 (-1 unseen), uv[n][2] int32 (optional). */
VFMREG_API int vfmreg_project_gather(vfmreg_ctx* ctx, const float* points, int64_t n, const vfmreg_camera* cams, int32_t n_cam,
                          const float* tokens, const int64_t* token_offsets, const uint8_t* images,
                          const int64_t* image_offsets, int32_t d, float* desc, int32_t* cam_of_point, int32_t* uv);

/* ---------------------------------------------------------------------------------------------
 * a1/a2  image features: ImageFeatureGenerator.get_image_features' GPU work (vfm_reg/image_features.py:67-77 transform,
 * :95-101 `self.model.model(x)` = FeatUp DINOv2 featurizer + ChannelNorm).  ViT-S/14 is what the reference loads
 * (feature_size 384, image_features.py:43-44); ViT-B/14 and ViT-L/14 are the same code with other dimensions.
 * GEMMs run on tcgen05 in bf16 with fp32 accumulation; the residual stream and all normalisations are fp32.
 * ------------------------------------------------------------------------------------------- */
typedef struct vfmreg_vit vfmreg_vit;

typedef struct {
  int32_t depth, width, heads, mlp_dim;  /* ViT-S 12/384/6/1536, ViT-B 12/768/12/3072, ViT-L 24/1024/16/4096 */
  int32_t patch;                         /* 14 */
  int32_t patch_h;                       /* patch rows after the resize: 16 (image_features.py:33) */
  int32_t channel_norm;                  /* 1: apply FeatUp's ChannelNorm (use_norm=True, image_features.py:41) */
  float ln_eps;                          /* 1e-6 */
  float cn_eps;                          /* 1e-4 */
  float mean[3], std[3];                 /* ImageNet statistics of the Normalize transform (image_features.py:76) */
} vfmreg_vit_config;

VFMREG_API int vfmreg_vit_create(vfmreg_ctx* ctx, const vfmreg_vit_config* cfg, vfmreg_vit** vit);
VFMREG_API void vfmreg_vit_destroy(vfmreg_vit* vit);
/* float32 host tensor by its dinov2-hub state-dict name (see vit.cu for the list); matrices become bf16 on the device */
VFMREG_API int vfmreg_vit_set_weight(vfmreg_vit* vit, const char* name, const float* host, int64_t count);
/* position embedding already interpolated to the patch grid: (1 + grid_h*grid_w, width) float32 host, CLS row first */
VFMREG_API int vfmreg_vit_set_pos_embed(vfmreg_vit* vit, int32_t grid_h, int32_t grid_w, const float* host);
/* patch grid create_transform_ would pick for an (img_h, img_w) image (image_features.py:67-69) */
VFMREG_API int vfmreg_vit_grid(const vfmreg_vit* vit, int32_t img_h, int32_t img_w, int32_t* grid_h, int32_t* grid_w);
/* 1 (default): from the second call with a given (b, img_h, img_w) the kernel sequence is replayed as one CUDA graph */
VFMREG_API int vfmreg_vit_set_graphs(vfmreg_vit* vit, int32_t on);
/* images: device uint8 (b, img_h, img_w, 3) RGB; tokens: device float32 (b, grid_h, grid_w, width) */
VFMREG_API int vfmreg_vit_forward(vfmreg_vit* vit, const uint8_t* images, int32_t b, int32_t img_h, int32_t img_w, float* tokens);

/* ---------------------------------------------------------------------------------------------
 * SURVEY 8f row 2 -- voxel down-sampling and the voxel hash map (the step BEFORE the path).
 *
 * vfmreg_voxel_downsample replaces kiss_icp::VoxelDownsample (core/Preprocessing.cpp:50-137; pybind
 * `_voxel_down_sample`, pybind/kiss_icp_pybind.cpp, python voxelization.voxel_down_sample): keeps the first point
 * (lowest index) of every voxel, voxel = (p / voxel_size).cast<int>() computed in double.
 *   points: device, n rows of `cols` elements of `elem_size` bytes (4 = float32, 8 = float64), xyz in columns 0..2
 *   keep_idx[n] (device): indices of the kept rows in ascending order; *count (device) = how many.
 *   The reference returns the rows in tsl::robin_map iteration order (unspecified); here they are in input order.
 * vfmreg_gather_rows copies the listed rows (row_bytes % 4 == 0): out[k] = rows[idx[k]], k < *count.
 * Error: VFMREG_ERR_ARG if a coordinate is not finite or |p / voxel_size| >= 2^20.
 * ------------------------------------------------------------------------------------------- */
VFMREG_API int vfmreg_voxel_downsample(vfmreg_ctx* ctx, const void* points, int64_t n, int32_t cols, int32_t elem_size,
                                       double voxel_size, int32_t* keep_idx, int32_t* count);
VFMREG_API int vfmreg_gather_rows(vfmreg_ctx* ctx, const void* rows, int32_t row_bytes, const int32_t* idx, const int32_t* count,
                                  int64_t max_rows, void* out);

/* VoxelHashMap (core/VoxelHashMap.hpp:40-75, VoxelHashMap.cpp:735-771 AddPoints, VoxelBlock::AddPoint): every voxel
 * keeps the first max_points_per_voxel points in insertion order.  build() replaces the content with the points of one
 * array (n x 3 float64, device); incremental add_points is done by the host wrapper, which rebuilds from
 * [kept points, new points] -- the same result as the reference's cumulative insertion.
 * points(): kept points grouped by voxel (ascending source index inside a voxel) + their index in the build array
 * (the reference's Pointcloud() order is the hash map's iteration order, i.e. unspecified).
 * nearest(): VoxelHashMap::GetCorrespondences' GetClosestNeighbor (VoxelHashMap.cpp:79-136): closest kept point among
 * the 27 voxels around the query, first minimum in (i, j, k, insertion) order; nn_idx = -1 unless its distance is
 * < max_dist; nn_d2 (may be NULL) = squared distance of the closest point (-1 if the 27 voxels are empty). */
typedef struct vfmreg_voxel_map vfmreg_voxel_map;
VFMREG_API int vfmreg_voxel_map_create(vfmreg_ctx* ctx, double voxel_size, int32_t max_points_per_voxel, vfmreg_voxel_map** map);
VFMREG_API void vfmreg_voxel_map_destroy(vfmreg_voxel_map* map);
VFMREG_API int vfmreg_voxel_map_build(vfmreg_ctx* ctx, vfmreg_voxel_map* map, const double* xyz, int64_t n);
VFMREG_API int64_t vfmreg_voxel_map_size(const vfmreg_voxel_map* map);
VFMREG_API int vfmreg_voxel_map_points(vfmreg_ctx* ctx, const vfmreg_voxel_map* map, double* xyz_out, int32_t* src_idx_out);
VFMREG_API int vfmreg_voxel_map_nearest(vfmreg_ctx* ctx, const vfmreg_voxel_map* map, const double* query, int64_t n,
                                        double max_dist, int32_t* nn_idx, double* nn_d2);

/* ---------------------------------------------------------------------------------------------
 * SURVEY 8f row 1 -- ICP refinement (the step AFTER the path).
 * Replaces kiss_icp::RegisterFrame (core/Registration.cpp:145-195; pybind `_register_frame`, python
 * registration.register_frame, called at registration_node.py:337-341): point-to-point Gauss-Newton with the
 * Geman-McClure weight kernel^2 / (kernel + r^2)^2 over the 27-voxel nearest neighbours, SE(3) exponential update,
 * stops when |dx| < 1e-4, when no correspondence is left, or after max_iterations (reference: 1000).
 *   frame: device (n x 3) float64; T0 / T_out: HOST row-major 4x4; returns T_icp * T0.
 *   iterations / correspondences (HOST, may be NULL): solves executed, correspondences of the last one.
 * An empty map returns T0 (Registration.cpp:150).  Sums are reduced in a fixed order (the reference's
 * tbb::parallel_reduce order is not), float64 throughout.
 * ------------------------------------------------------------------------------------------- */
VFMREG_API int vfmreg_register_frame(vfmreg_ctx* ctx, const vfmreg_voxel_map* map, const double* frame, int64_t n, const double* T0,
                                     double max_correspondence_distance, double kernel, int32_t max_iterations, double* T_out,
                                     int32_t* iterations, int32_t* correspondences);

/* The descriptor-carrying overload of RegisterFrame ("VFM-ICP", core/Registration.cpp:197-382; reached through
 * register_frame with (N, 3 + D) frames, kiss_icp/registration.py:41-62).  vfm_src / vfm_tgt (device, k x 3 float64) are the
 * descriptor correspondences the reference computes first (VoxelDownsample(source, 5.0) -> GetVFMCorrespondences(.., 0.8);
 * source points BEFORE the initial guess is applied, map points).  Loop 1: Gauss-Newton on that fixed list, pruned after
 * every update to |d - median| < 1.5 * 1.4826 * MAD, until the mean distance moves by < 0.01 m.  Loop 2: the vanilla ICP of
 * vfmreg_register_frame with the remaining iteration budget.  vfm_iterations / iterations / vfm_kept (HOST, may be NULL):
 * iterations of loop 1, of loop 2, correspondences left after the pruning. */
VFMREG_API int vfmreg_register_frame_vfm(vfmreg_ctx* ctx, const vfmreg_voxel_map* map, const double* frame, int64_t n,
                                         const double* vfm_src, const double* vfm_tgt, int64_t k, const double* T0,
                                         double max_correspondence_distance, double kernel, int32_t max_iterations, double* T_out,
                                         int32_t* vfm_iterations, int32_t* iterations, int32_t* vfm_kept);

/* ---------------------------------------------------------------------------------------------
 * SURVEY 8f row 4 -- the hypothesis score of Open3D 0.18's registration_ransac_based_on_correspondence, the solver the
 * reference actually calls (registration_node.py:312-327): every hypothesis (3 sampled correspondences -> rigid fit, as in
 * vfmreg_ransac) transforms the WHOLE source cloud `src_all` and is scored against the nearest target point of every
 * transformed point: fitness = #(nearest neighbour closer than max_dist) / n_src, rmse over those; best = larger fitness,
 * then smaller rmse, then lower hypothesis id; no refit.  With the reference's literal max_dist = 10000 the winner is the
 * hypothesis with the smallest scan -> map chamfer RMSE.  Open3D's source is not in the reference tree: this follows its
 * published algorithm as recalled in SURVEY.md A.8 ("parity unpinned").
 *   vfmreg_kdtree_create   balanced k-d tree over the target cloud (HOST float64 n x 3; built on the host, kept on the device)
 *   vfmreg_kdtree_nearest  exact nearest neighbour of device queries (n x 3 float64): nn_idx in the caller's order (-1 when
 *                          nothing is closer than max_dist), nn_d2 (may be NULL) its squared distance
 *   vfmreg_ransac_nn_all   src_all (n_src x 3 float64 device); src_xyz / tgt_xyz / corr / count / sample_idx / seed as in
 *                          vfmreg_ransac (the clouds the correspondences index); outputs (device): T[16], inliers[n_hyp] and
 *                          sum_d2[n_hyp] (optional), stats[4] = {best hypothesis or -1, its inlier count, the bit pattern of
 *                          its sum of squared distances (a double), unused}.
 * ------------------------------------------------------------------------------------------- */
typedef struct vfmreg_kdtree vfmreg_kdtree;
VFMREG_API int vfmreg_kdtree_create(vfmreg_ctx* ctx, const double* xyz_host, int64_t n, vfmreg_kdtree** tree);
VFMREG_API void vfmreg_kdtree_destroy(vfmreg_kdtree* tree);
VFMREG_API int vfmreg_kdtree_nearest(vfmreg_ctx* ctx, const vfmreg_kdtree* tree, const double* queries, int64_t n, double max_dist,
                                     int32_t* nn_idx, double* nn_d2);
VFMREG_API int vfmreg_ransac_nn_all(vfmreg_ctx* ctx, const vfmreg_kdtree* tree, const double* src_all, int64_t n_src, const void* src_xyz,
                                    const void* tgt_xyz, int xyz_f64, const int32_t* corr, const int32_t* count, int32_t max_corr,
                                    const int32_t* sample_idx, int32_t n_hyp, uint64_t seed, double max_dist, double* T,
                                    int32_t* inliers, double* sum_d2, int64_t* stats);

/* ---------------------------------------------------------------------------------------------
 * SURVEY 8f row 4, second half -- the TEASER++ solve of registration_node.py:91-131 (teaserpp_python.RobustRegistrationSolver
 * with the parameters set there: cbar2 = 1, noise_bound = 0.2, no scale estimation, PMC_EXACT inlier selection, CHAIN TIM graph,
 * GNC-TLS rotation with factor 1.4 / 10000 iterations / cost threshold 1e-16).  TEASER++ is not in the reference tree: this is
 * its published algorithm ("parity unpinned"): compatibility graph of the translation-invariant measurements (GPU), exact
 * maximum clique (host branch and bound, `max_clique_nodes` search nodes at most: 0 = 20 M), GNC-TLS rotation on the clique's
 * chain of TIMs, component-wise TLS translation by adaptive voting.
 *   src_xyz / tgt_xyz  HOST float64 (k x 3): the correspondences' source and target points (solver.solve(src, tgt))
 *   T                  HOST double[16], row-major 4 x 4 (identity when k < 2 or the clique has fewer than 2 vertices)
 *   clique             HOST int32[k] (optional): the clique's correspondence indices, ascending
 *   stats              HOST int32[5] (optional): clique size, 1 if the clique search finished (exact), GNC iterations,
 *                      rotation inliers (TIMs with weight >= 0.5), translation inliers (minimum over the three axes)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  double noise_bound;       /* solver_params.noise_bound (0.2: "should be similar to the voxel size") */
  double cbar2;             /* solver_params.cbar2 (1.0) */
  double gnc_factor;        /* rotation_gnc_factor (1.4) */
  double cost_threshold;    /* rotation_cost_threshold (1e-16) */
  int32_t max_iterations;   /* rotation_max_iterations (10000) */
  int32_t reserved;
  int64_t max_clique_nodes; /* 0 = default */
} vfmreg_teaser_params;
VFMREG_API int vfmreg_teaser_solve(vfmreg_ctx* ctx, const double* src_xyz, const double* tgt_xyz, int64_t k, const vfmreg_teaser_params* p,
                                   double* T, int32_t* clique, int32_t* stats);

#ifdef __cplusplus
}
#endif
#endif /* VFMREG_B200_H */
