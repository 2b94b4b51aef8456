// Shared internals of libvfmreg_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/vfmreg_b200.h"

struct vfmreg_ctx;

namespace vfm {

void set_error(const char* fmt, ...);

#define VFM_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      vfm::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return VFMREG_ERR_CUDA;                                                               \
    }                                                                                       \
  } while (0)

#define VFM_CHECK_ARG(cond, ...)   \
  do {                             \
    if (!(cond)) {                 \
      vfm::set_error(__VA_ARGS__); \
      return VFMREG_ERR_ARG;       \
    }                              \
  } while (0)

#define VFM_TRY(expr)                 \
  do {                                \
    int rc__ = (expr);                \
    if (rc__ != VFMREG_OK) return rc__; \
  } while (0)

// GROUP_MATCH_PRUNED: the candidate-search launches whose row count lives on the device (pruned reverse search)
// groups >= 5: the small kernels of a pair, stage by stage (tools/stage_times.py)
enum { GROUP_MATCH = 0, GROUP_RANSAC = 1, GROUP_PROJECT = 2, GROUP_VIT = 3, GROUP_MATCH_PRUNED = 4, GROUP_NORMALIZE = 5,
       GROUP_RERANK = 6, GROUP_FILTER = 7, GROUP_GATHER = 8, GROUP_MUTUAL = 9, GROUP_KABSCH = 10, GROUP_FINALIZE = 11,
       NUM_GROUPS = 12 };

// times everything enqueued on the context's stream inside a scope as one interval of `group`
struct GroupScope {
  vfmreg_ctx* ctx;
  int group, launches;
  GroupScope(vfmreg_ctx* c, int g, int n);
  ~GroupScope();
};

// Bump allocator over one cudaMalloc'd slab.  Scratch is carved per API call and reused by the next call on the
// same stream (stream order makes that safe); growing the slab synchronises the stream first.
struct Arena {
  char* base = nullptr;
  size_t cap = 0;
  size_t off = 0;
  size_t limit = 0;   // end of the region the current carve may use (0 = the whole slab): a batch lane must not run into the next
};

// A descriptor set ready for the search: renormalised fp32 rows (rows x dp, zero padded), and on the tensor-core path
// their fp16 copy and per-row non-zero flags.
struct Prepared {
  const float* f32 = nullptr;
  const uint16_t* f16 = nullptr;
  const uint8_t* nz = nullptr;
  int64_t rows = 0;
  int dp = 0;
  bool tc = false;
};

}  // namespace vfm

namespace vfm {
struct HostCopy;   // hostcopy.cu: staging ring + copy threads for pageable host sources
}

struct vfmreg_ctx {
  int device = 0;
  vfm::HostCopy* hostcopy = nullptr;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  vfm::Arena arena;
  int64_t launches = 0;
  // optional per-group device timing (CUDA events on the context's stream)
  int timing = 0;
  // a ring of event pairs per group, so that timing a launch never makes the host wait for the previous one
  static constexpr int EV_RING = 64;
  cudaEvent_t ev0[vfm::NUM_GROUPS][EV_RING] = {}, ev1[vfm::NUM_GROUPS][EV_RING] = {};
  int ev_head[vfm::NUM_GROUPS] = {};   // next ring slot to record into
  float group_ms[vfm::NUM_GROUPS] = {};
  int group_launches[vfm::NUM_GROUPS] = {};
  int group_pending[vfm::NUM_GROUPS] = {};   // recorded intervals not yet folded into group_ms (<= EV_RING)
  // pinned staging for small results
  void* pinned = nullptr;
  size_t pinned_cap = 0;
  // persistent device buffers for register_host (inputs), grown on demand
  char* hbuf = nullptr;
  size_t hbuf_cap = 0;
  // register_batch_host: second stream + events for the double-buffered H2D / compute pipeline
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_ready[2] = {}, ev_consumed[2] = {}, ev_target_free[2] = {};
  // batch entry points: a second compute lane.  Consecutive pairs alternate between the caller's stream and this one, so
  // the latency-bound small kernels of one pair (filters, re-rank, RANSAC) run beside the other pair's match kernel.
  static constexpr int MAX_LANES = 16;
  int lanes = 8;                                   // vfmreg_set_lanes: pairs per group
  cudaStream_t lane_stream[MAX_LANES] = {};        // [0] unused: lane 0 is the caller's stream
  cudaEvent_t ev_fork = nullptr, ev_join[MAX_LANES] = {};
  // batch mode: high-priority streams for the candidate-search kernels; null outside a batch
  static constexpr int MATCH_EVENTS = 128;
  cudaStream_t match_stream = nullptr;         // the batch's search stream while a batch is being enqueued, else null
  cudaStream_t match_stream_owned = nullptr;
  cudaEvent_t match_ev[MATCH_EVENTS] = {};
  int match_ev_head = 0;
  // cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remembered per context, not per process
  bool tc_attr_set = false;
  int gemm_clusters = 0;           // resident clusters of vit_gemm_kernel on this device (0 = not asked yet)
  uint64_t gemm_attr_mask = 0;     // vit_gemm_kernel<EPI, BN> instantiations already opted in
  size_t attention_tc_smem_attr = 0;   // same for attention_tc_kernel
  size_t attention_smem_attr = 0;  // largest dynamic shared memory size attention_kernel was opted in for
  // events that order the pairs of a batch after the preparation of the map they share (grown on demand)
  cudaEvent_t* map_ev = nullptr;
  int map_ev_cap = 0;
};

namespace vfm {

// hostcopy.cu: host -> device copy of a caller-owned buffer on `stream`; pageable sources are staged through pinned memory by a
// pool of copy threads (the source may be reused when the call returns)
int h2d_copy(vfmreg_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t stream);
void hostcopy_destroy(vfmreg_ctx* ctx);
int arena_reserve(vfmreg_ctx* ctx, size_t bytes);  // make sure the slab holds `bytes` in total (call before carving)
inline void arena_reset(vfmreg_ctx* ctx) { ctx->arena.off = 0; }
template <typename T>
inline T* arena_take(vfmreg_ctx* ctx, size_t count) {
  size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
  if (ctx->arena.off + bytes > (ctx->arena.limit ? ctx->arena.limit : ctx->arena.cap)) return nullptr;
  T* p = reinterpret_cast<T*>(ctx->arena.base + ctx->arena.off);
  ctx->arena.off += bytes;
  return p;
}
inline size_t arena_bytes(size_t count, size_t elem) { return (count * elem + 255) & ~size_t(255); }

// NVTX ranges per stage (visible in nsys / ncu --nvtx; free when no tool is attached)
void nvtx_push(const char* name);
void nvtx_pop();
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtx_push(name); }
  ~NvtxRange() { nvtx_pop(); }
};

void group_begin(vfmreg_ctx* ctx, int group);
void group_end(vfmreg_ctx* ctx, int group, int n_launches);
inline GroupScope::GroupScope(vfmreg_ctx* c, int g, int n) : ctx(c), group(g), launches(n) { group_begin(c, g); }
inline GroupScope::~GroupScope() { group_end(ctx, group, launches); }

inline int launch_check(vfmreg_ctx* ctx, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("kernel launch %s failed: %s", what, cudaGetErrorString(e));
    return VFMREG_ERR_CUDA;
  }
  ctx->launches += 1;
  return VFMREG_OK;
}

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- stage entry points (each file implements its own) -------------------------------------------------
// normalize.cu: rows of x (n x d) -> y (n x dp) zero padded, L2-renormalised when `normalize`
// optional yh: fp16 copy (n x dp halves); optional nz: 1 per row unless the row is all-zero
int normalize_rows(vfmreg_ctx* ctx, const float* x, int64_t n, int d, int dp, int normalize, float* y, void* yh = nullptr,
                   uint8_t* nz = nullptr);
// match_simt.cu: exact fp32 top-2 of a (n x dp) against b (m x dp)
int match_simt(vfmreg_ctx* ctx, const float* a, int64_t n, const float* b, int64_t m, int dp, int32_t* idx,
               float* best, float* sec);
size_t match_simt_scratch(vfmreg_ctx* ctx, int64_t n, int64_t m);
// match_tc.cu: tcgen05 fp16 candidate search + exact fp32 re-rank; same results as match_simt (renormalised inputs only)
int match_tc(vfmreg_ctx* ctx, const float* a32, const void* a16, const uint8_t* nz_a, int64_t n, const float* b32,
             const void* b16, int64_t m, int dp, int32_t* idx, float* best, float* sec, const int* n_dev = nullptr,
             const float* seed = nullptr, float floor = NAN);
size_t match_tc_scratch(vfmreg_ctx* ctx, int64_t n, int64_t m, bool dynamic = false);
// the two halves of match_tc: candidate search, then re-rank (see match_tc.cu)
struct TcPending {
  const float *a32 = nullptr, *b32 = nullptr;
  const uint8_t* nz_a = nullptr;
  int64_t n = 0, m = 0;
  int dp = 0, slots = 0, top1 = 0, floor_mode = 0;
  const int* n_dev = nullptr;
  float* cand_v = nullptr;
  int *cand_i = nullptr, *cand_n = nullptr;
  void* slot_top2 = nullptr;
  int32_t* idx = nullptr;
  float *best = nullptr, *sec = nullptr;
  cudaEvent_t done = nullptr;   // recorded behind the search kernel on the batch's search stream (null outside a batch)
};
int match_tc_begin(vfmreg_ctx* ctx, const float* a32, const void* a16, const uint8_t* nz_a, int64_t n, const float* b32,
                   const void* b16, int64_t m, int dp, int32_t* idx, float* best, float* sec, const int* n_dev, const float* seed,
                   float floor, TcPending* pending);
int match_tc_finish(vfmreg_ctx* ctx, const TcPending& pending, cudaEvent_t wait_for);
// gather the fp32 / fp16 rows and non-zero flags of b listed in column `col` of the (count, 2) list `pairs`
int gather_rows(vfmreg_ctx* ctx, const int32_t* pairs, const int32_t* count, int64_t max_rows, int col, int dp, const float* b32,
                const void* b16, const uint8_t* nzb, float* o32, void* o16, uint8_t* onz, const float* sim = nullptr,
                float* osim = nullptr);   // osim[k] = sim[pairs[2k]] (the forward score of the listed pair)
// filter.cu
int filter_corr(vfmreg_ctx* ctx, const int32_t* idx01, const float* sim01, const float* sec01, const int32_t* idx10,
                int64_t n, float min_cos, float ratio, int mutual, int32_t* corr, int32_t* count);
// keep the n_keep queries with the largest similarity (= smallest L2 distance of unit vectors), ties towards the lowest
// index, pairs in query order; dist (optional) = sqrt(2 - 2 s + 1e-6) of the kept pairs
int select_top(vfmreg_ctx* ctx, const int32_t* idx01, const float* sim01, int64_t n, int64_t n_keep, int32_t* corr, float* dist,
               int32_t* count);
int l2_from_sim(vfmreg_ctx* ctx, const int32_t* idx01, const float* sim01, int64_t n, float* dist);
// second half of the pruned mutual check: keep cand[k] = (i, j) iff back[k] == i (back = nearest query of map row j)
int filter_mutual_list(vfmreg_ctx* ctx, const int32_t* cand, const int32_t* cand_count, const int32_t* back, int64_t max_rows,
                       int32_t* corr, int32_t* count);
// ransac.cu
size_t ransac_scratch(int32_t max_corr, int32_t n_hyp);
int ransac_solve(vfmreg_ctx* ctx, const void* src_xyz, const void* tgt_xyz, int xyz_f64, const int32_t* corr,
                 const int32_t* count, int32_t max_corr, const int32_t* sample_idx, int32_t n_hyp, uint64_t seed,
                 double thresh, int refit, double* T, int32_t* counts, int64_t* sumq, uint8_t* mask, int64_t* stats);

}  // namespace vfm
