// extern "C" surface of libvfmreg_b200.so (see include/vfmreg_b200.h).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace vfm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int arena_reserve(vfmreg_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->arena.cap) return VFMREG_OK;
  VFM_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->arena.base) VFM_CUDA(cudaFree(ctx->arena.base));
  ctx->arena.base = nullptr;
  ctx->arena.cap = 0;
  size_t want = bytes + (bytes >> 2) + (1 << 20);
  cudaError_t e = cudaMalloc(&ctx->arena.base, want);
  if (e != cudaSuccess) {
    set_error("scratch arena: cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    return VFMREG_ERR_ALLOC;
  }
  ctx->arena.cap = want;
  ctx->arena.off = 0;
  return VFMREG_OK;
}

// fold the oldest `n` pending intervals of a group into its total (waits for them if they have not completed yet)
static void group_fold(vfmreg_ctx* ctx, int group, int n) {
  const int R = vfmreg_ctx::EV_RING;
  for (; n > 0 && ctx->group_pending[group] > 0; --n) {
    const int slot = ((ctx->ev_head[group] - ctx->group_pending[group]) % R + R) % R;
    float ms = 0.f;
    if (cudaEventSynchronize(ctx->ev1[group][slot]) == cudaSuccess &&
        cudaEventElapsedTime(&ms, ctx->ev0[group][slot], ctx->ev1[group][slot]) == cudaSuccess)
      ctx->group_ms[group] += ms;
    ctx->group_pending[group] -= 1;
  }
}

void group_begin(vfmreg_ctx* ctx, int group) {
  if (!ctx->timing) return;
  if (ctx->group_pending[group] == vfmreg_ctx::EV_RING) group_fold(ctx, group, 1);  // ring full: reuse the oldest slot
  cudaEventRecord(ctx->ev0[group][ctx->ev_head[group]], ctx->stream);
}

void group_end(vfmreg_ctx* ctx, int group, int n_launches) {
  if (!ctx->timing) return;
  cudaEventRecord(ctx->ev1[group][ctx->ev_head[group]], ctx->stream);
  ctx->ev_head[group] = (ctx->ev_head[group] + 1) % vfmreg_ctx::EV_RING;
  ctx->group_pending[group] += 1;
  ctx->group_launches[group] += n_launches;
}

static int ensure_pinned(vfmreg_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->pinned_cap) return VFMREG_OK;
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  ctx->pinned = nullptr;
  ctx->pinned_cap = 0;
  VFM_CUDA(cudaMallocHost(&ctx->pinned, bytes));
  ctx->pinned_cap = bytes;
  return VFMREG_OK;
}

static inline int round_up(int x, int q) { return (x + q - 1) / q * q; }

static bool use_tc(uint32_t flags, int d) {
  const uint32_t algo = flags & VFMREG_ALGO_MASK;
  // the tcgen05 path's error bound assumes unit-norm rows of at most 1024 dimensions; un-normalised or wider searches
  // stay on the exact fp32 kernel
  return (flags & VFMREG_NORMALIZE) && algo != VFMREG_ALGO_SIMT && d <= 1024;
}

static inline int padded_dim(int d, uint32_t flags) { return round_up(d, use_tc(flags, d) ? 64 : 16); }

// scratch needed by match_nn_impl beyond the caller-visible outputs
static size_t match_scratch(vfmreg_ctx* ctx, int64_t n, int64_t m, int d, uint32_t flags) {
  const int dp = padded_dim(d, flags);
  const bool mutual = (flags & VFMREG_MUTUAL) != 0;
  size_t s = arena_bytes((size_t)n * dp, 4) + arena_bytes((size_t)m * dp, 4);
  if (use_tc(flags, d)) {
    s += arena_bytes((size_t)n * dp, 2) + arena_bytes((size_t)m * dp, 2) + arena_bytes(n, 1) + arena_bytes(m, 1);
    s += match_tc_scratch(ctx, n, m);
    if (mutual) s += match_tc_scratch(ctx, m, n);
  } else {
    s += match_simt_scratch(ctx, n, m);
    if (mutual) s += match_simt_scratch(ctx, m, n);
  }
  return s;
}

// VFMREG_FULL_MUTUAL=1 keeps the full reverse search inside register() (A/B comparison; results are identical)
static bool g_full_mutual = [] { const char* e = getenv("VFMREG_FULL_MUTUAL"); return e && e[0] == '1'; }();

// The reverse search of the mutual check can be restricted to the map rows that some gated query points at when the
// tensor-core path runs and the query side is the smaller one (the pruned search is (<= n) x n instead of m x n).
static bool prune_mutual(int64_t n, int64_t m, int d, uint32_t flags) {
  return use_tc(flags, d) && (flags & VFMREG_MUTUAL) && n <= m && !g_full_mutual;
}

static size_t pruned_scratch(vfmreg_ctx* ctx, int64_t n, int d, uint32_t flags) {
  const int dp = padded_dim(d, flags);
  return arena_bytes((size_t)n * 2, 4) + arena_bytes(1, 4) + arena_bytes((size_t)n * dp, 4) + arena_bytes((size_t)n * dp, 2) +
         arena_bytes(n, 1) + arena_bytes(n, 4) * 2 + match_tc_scratch(ctx, n, n, true);
}

static int match_nn_impl(vfmreg_ctx* ctx, const float* a, int64_t n, const float* b, int64_t m, int32_t d, uint32_t flags,
                         int32_t* idx01, float* sim01, float* sec01, int32_t* idx10, float* sim10, float* sec10,
                         const vfmreg_register_params* prune = nullptr, int32_t* corr = nullptr, int32_t* count = nullptr) {
  const bool tc = use_tc(flags, d);
  const int dp = padded_dim(d, flags);
  float* an = arena_take<float>(ctx, (size_t)n * dp);
  float* bn = arena_take<float>(ctx, (size_t)m * dp);
  uint16_t *ah = nullptr, *bh = nullptr;
  uint8_t *nza = nullptr, *nzb = nullptr;
  if (tc) {
    ah = arena_take<uint16_t>(ctx, (size_t)n * dp);
    bh = arena_take<uint16_t>(ctx, (size_t)m * dp);
    nza = arena_take<uint8_t>(ctx, n);
    nzb = arena_take<uint8_t>(ctx, m);
  }
  if (!an || !bn || (tc && (!ah || !bh || !nza || !nzb))) {
    set_error("match_nn: scratch arena too small");
    return VFMREG_ERR_ALLOC;
  }
  const int norm = (flags & VFMREG_NORMALIZE) != 0;
  VFM_TRY(normalize_rows(ctx, a, n, d, dp, norm, an, ah, nza));
  VFM_TRY(normalize_rows(ctx, b, m, d, dp, norm, bn, bh, nzb));
  if (prune) {
    // register(): forward search, gate (cosine / ratio) -> candidate list (i, j) in query order, reverse search over the
    // listed map rows only, then keep the candidates whose map row points back at them
    int32_t* cand = arena_take<int32_t>(ctx, (size_t)n * 2);
    int32_t* cand_count = arena_take<int32_t>(ctx, 1);
    float* sel32 = arena_take<float>(ctx, (size_t)n * dp);
    uint16_t* sel16 = arena_take<uint16_t>(ctx, (size_t)n * dp);
    uint8_t* selnz = arena_take<uint8_t>(ctx, n);
    int32_t* back = arena_take<int32_t>(ctx, n);
    float* selsim = arena_take<float>(ctx, n);
    if (!cand || !cand_count || !sel32 || !sel16 || !selnz || !back || !selsim) {
      set_error("register: scratch arena too small");
      return VFMREG_ERR_ALLOC;
    }
    VFM_TRY(match_tc(ctx, an, ah, nza, n, bn, bh, m, dp, idx01, sim01, sec01));
    VFM_TRY(filter_corr(ctx, idx01, sim01, sec01, nullptr, n, prune->min_cos, prune->ratio, 0, cand, cand_count));
    VFM_TRY(gather_rows(ctx, cand, cand_count, n, 1, dp, bn, bh, nzb, sel32, sel16, selnz, sim01, selsim));
    // <b_j, a_i> has the same canonical value as <a_i, b_j> = sim01[i]: a lower bound of row j's best that starts the
    // candidate recording near the answer
    VFM_TRY(match_tc(ctx, sel32, sel16, selnz, n, an, ah, n, dp, back, nullptr, nullptr, cand_count, selsim));
    return filter_mutual_list(ctx, cand, cand_count, back, n, corr, count);
  }
  if (flags & VFMREG_MUTUAL) VFM_CHECK_ARG(idx10, "match_nn: VFMREG_MUTUAL needs idx10");
  if (tc) {
    VFM_TRY(match_tc(ctx, an, ah, nza, n, bn, bh, m, dp, idx01, sim01, sec01));
    if (flags & VFMREG_MUTUAL) VFM_TRY(match_tc(ctx, bn, bh, nzb, m, an, ah, n, dp, idx10, sim10, sec10));
  } else {
    VFM_TRY(match_simt(ctx, an, n, bn, m, dp, idx01, sim01, sec01));
    if (flags & VFMREG_MUTUAL) VFM_TRY(match_simt(ctx, bn, m, an, n, dp, idx10, sim10, sec10));
  }
  return VFMREG_OK;
}

}  // namespace vfm

using namespace vfm;

extern "C" {

int vfmreg_version(void) { return VFMREG_VERSION; }

const char* vfmreg_last_error(void) { return g_err; }

int vfmreg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int vfmreg_create(int device, vfmreg_ctx** out) {
  VFM_CHECK_ARG(out, "vfmreg_create: null out pointer");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device available (%s); this library has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return VFMREG_ERR_NOGPU;
  }
  VFM_CHECK_ARG(device >= 0 && device < n, "vfmreg_create: device %d out of range (count %d)", device, n);
  cudaDeviceProp prop;
  VFM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libvfmreg_b200 carries sm_100a code only", device, prop.major, prop.minor);
    return VFMREG_ERR_NOGPU;
  }
  VFM_CUDA(cudaSetDevice(device));
  vfmreg_ctx* ctx = new vfmreg_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if (const char* e = getenv("VFMREG_LANES")) {   // tuning aid; vfmreg_set_lanes is the API
    const int l = atoi(e);
    if (l >= 1 && l <= vfmreg_ctx::MAX_LANES) ctx->lanes = l;
  }
  for (int g = 0; g < NUM_GROUPS; ++g) {
    for (int r = 0; r < vfmreg_ctx::EV_RING; ++r) {
      cudaEventCreate(&ctx->ev0[g][r]);
      cudaEventCreate(&ctx->ev1[g][r]);
    }
  }
  *out = ctx;
  return VFMREG_OK;
}

void vfmreg_destroy(vfmreg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->arena.base) cudaFree(ctx->arena.base);
  if (ctx->hbuf) cudaFree(ctx->hbuf);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i < 2; ++i) {
      cudaEventDestroy(ctx->ev_ready[i]);
      cudaEventDestroy(ctx->ev_consumed[i]);
    }
  }
  for (int l = 1; l < vfmreg_ctx::MAX_LANES; ++l) {
    if (!ctx->lane_stream[l]) continue;
    cudaStreamSynchronize(ctx->lane_stream[l]);
    cudaStreamDestroy(ctx->lane_stream[l]);
    cudaEventDestroy(ctx->ev_join[l]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  for (int i = 0; i < 2; ++i)
    if (ctx->match_stream_owned[i]) {
      cudaStreamSynchronize(ctx->match_stream_owned[i]);
      cudaStreamDestroy(ctx->match_stream_owned[i]);
    }
  for (int i = 0; i < vfmreg_ctx::MATCH_EVENTS; ++i)
    if (ctx->match_ev[i]) cudaEventDestroy(ctx->match_ev[i]);
  for (int g = 0; g < NUM_GROUPS; ++g) {
    for (int r = 0; r < vfmreg_ctx::EV_RING; ++r) {
      cudaEventDestroy(ctx->ev0[g][r]);
      cudaEventDestroy(ctx->ev1[g][r]);
    }
  }
  delete ctx;
}

int vfmreg_set_stream(vfmreg_ctx* ctx, void* stream) {
  VFM_CHECK_ARG(ctx, "null context");
  if ((cudaStream_t)stream != ctx->stream) {
    VFM_CUDA(cudaSetDevice(ctx->device));
    VFM_CUDA(cudaStreamSynchronize(ctx->stream));  // scratch is reused in stream order
    ctx->stream = (cudaStream_t)stream;
  }
  return VFMREG_OK;
}

int vfmreg_sync(vfmreg_ctx* ctx) {
  VFM_CHECK_ARG(ctx, "null context");
  VFM_CUDA(cudaSetDevice(ctx->device));
  VFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return VFMREG_OK;
}

int64_t vfmreg_kernel_launches(const vfmreg_ctx* ctx) { return ctx ? ctx->launches : 0; }

int vfmreg_set_lanes(vfmreg_ctx* ctx, int lanes) {
  VFM_CHECK_ARG(ctx, "null context");
  VFM_CHECK_ARG(lanes >= 1 && lanes <= vfmreg_ctx::MAX_LANES, "set_lanes: %d not in [1, %d]", lanes, vfmreg_ctx::MAX_LANES);
  ctx->lanes = lanes;
  return VFMREG_OK;
}

int vfmreg_enable_timing(vfmreg_ctx* ctx, int on) {
  VFM_CHECK_ARG(ctx, "null context");
  ctx->timing = on;
  for (int g = 0; g < NUM_GROUPS; ++g) {
    ctx->group_ms[g] = 0.f;
    ctx->group_launches[g] = 0;
    ctx->group_pending[g] = 0;
    ctx->ev_head[g] = 0;
  }
  return VFMREG_OK;
}

int vfmreg_group_time_ms(vfmreg_ctx* ctx, int group, float* ms_total, int* launches) {
  VFM_CHECK_ARG(ctx && group >= 0 && group < NUM_GROUPS, "bad group");
  group_fold(ctx, group, ctx->group_pending[group]);
  if (ms_total) *ms_total = ctx->group_ms[group];
  if (launches) *launches = ctx->group_launches[group];
  return VFMREG_OK;
}

int vfmreg_match_nn(vfmreg_ctx* ctx, const float* a, int64_t n, const float* b, int64_t m, int32_t d, uint32_t flags,
                    int32_t* idx01, float* sim01, float* sec01, int32_t* idx10, float* sim10, float* sec10) {
  VFM_CHECK_ARG(ctx, "null context");
  VFM_CHECK_ARG(a && b && idx01, "match_nn: null pointer");
  VFM_CHECK_ARG(n > 0 && m > 0 && d > 0, "match_nn: empty input (n=%lld m=%lld d=%d)", (long long)n, (long long)m, d);
  VFM_CUDA(cudaSetDevice(ctx->device));
  arena_reset(ctx);
  VFM_TRY(arena_reserve(ctx, match_scratch(ctx, n, m, d, flags)));
  return match_nn_impl(ctx, a, n, b, m, d, flags, idx01, sim01, sec01, idx10, sim10, sec10);
}

int vfmreg_filter_correspondences(vfmreg_ctx* ctx, const int32_t* idx01, const float* sim01, const float* sec01,
                                  const int32_t* idx10, int64_t n, float min_cos, float ratio, int mutual,
                                  int32_t* corr, int32_t* count) {
  VFM_CHECK_ARG(ctx && idx01 && sim01 && corr && count, "filter_correspondences: null pointer");
  VFM_CHECK_ARG(n >= 0, "filter_correspondences: negative n");
  VFM_CUDA(cudaSetDevice(ctx->device));
  return filter_corr(ctx, idx01, sim01, sec01, idx10, n, min_cos, ratio, mutual, corr, count);
}

int vfmreg_ransac(vfmreg_ctx* ctx, const void* src_xyz, const void* tgt_xyz, int xyz_f64, const int32_t* corr,
                  const int32_t* count, int32_t max_corr, const int32_t* sample_idx, int32_t n_hyp, uint64_t seed,
                  double thresh, int refit, double* T, int32_t* counts, int64_t* sumq, uint8_t* mask, int64_t* stats) {
  VFM_CHECK_ARG(ctx && src_xyz && tgt_xyz && corr && count && T && stats, "ransac: null pointer");
  VFM_CUDA(cudaSetDevice(ctx->device));
  arena_reset(ctx);
  VFM_TRY(arena_reserve(ctx, ransac_scratch(max_corr, n_hyp)));
  return ransac_solve(ctx, src_xyz, tgt_xyz, xyz_f64, corr, count, max_corr, sample_idx, n_hyp, seed, thresh, refit, T,
                      counts, sumq, mask, stats);
}

struct RegOut {
  double* T;        // device, 16
  int64_t* stats;   // device, 8 (stats[0..3] + correspondence count as int32 at [4])
  int32_t* corr;    // device, n x 2
  uint8_t* mask;    // device, n
};

static size_t register_scratch(vfmreg_ctx* ctx, int64_t n, int64_t m, int32_t d, const vfmreg_register_params* p) {
  const bool mutual = (p->flags & VFMREG_MUTUAL) != 0;
  (void)mutual;
  return match_scratch(ctx, n, m, d, p->flags) + (prune_mutual(n, m, d, p->flags) ? pruned_scratch(ctx, n, d, p->flags) : 0) +
         ransac_scratch((int32_t)n, p->n_hyp) + arena_bytes(n, 4) * 3 + arena_bytes(m, 4) +
         arena_bytes((size_t)n * 2, 4) + arena_bytes(n, 1) + arena_bytes(16, 8) + arena_bytes(8, 8) + 4096;
}

// Enqueue the whole path on ctx->stream (no host synchronisation).  The arena must already be reserved.
static int register_enqueue(vfmreg_ctx* ctx, const float* src_xyz, const float* tgt_xyz, const float* src_feats,
                            const float* tgt_feats, int64_t n, int64_t m, int32_t d, const vfmreg_register_params* p,
                            const int32_t* sample_idx, const RegOut& out) {
  const bool mutual = (p->flags & VFMREG_MUTUAL) != 0;
  const bool use_ratio = !(p->ratio != p->ratio);
  int32_t* idx01 = arena_take<int32_t>(ctx, n);
  float* sim01 = arena_take<float>(ctx, n);
  float* sec01 = arena_take<float>(ctx, n);
  const bool pruned = prune_mutual(n, m, d, p->flags);
  int32_t* idx10 = (mutual && !pruned) ? arena_take<int32_t>(ctx, m) : nullptr;
  if (!idx01 || !sim01 || !sec01 || (mutual && !pruned && !idx10)) {
    set_error("register: scratch arena too small");
    return VFMREG_ERR_ALLOC;
  }
  int32_t* count = reinterpret_cast<int32_t*>(out.stats + 4);
  if (pruned) {
    VFM_TRY(match_nn_impl(ctx, src_feats, n, tgt_feats, m, d, p->flags, idx01, sim01, use_ratio ? sec01 : nullptr, nullptr,
                          nullptr, nullptr, p, out.corr, count));
  } else {
    VFM_TRY(match_nn_impl(ctx, src_feats, n, tgt_feats, m, d, p->flags, idx01, sim01, use_ratio ? sec01 : nullptr, idx10,
                          nullptr, nullptr));
    VFM_TRY(filter_corr(ctx, idx01, sim01, sec01, idx10, n, p->min_cos, p->ratio, mutual, out.corr, count));
  }
  return ransac_solve(ctx, src_xyz, tgt_xyz, 0, out.corr, count, (int32_t)n, sample_idx, p->n_hyp, p->seed, p->inlier_thresh,
                      p->refit, out.T, nullptr, nullptr, out.mask, out.stats);
}

static void fill_result(vfmreg_register_result* result, const char* pin, double thresh) {
  memcpy(result->T, pin, 16 * sizeof(double));
  const int64_t* st = reinterpret_cast<const int64_t*>(pin + 128);
  result->best_hyp = st[0];
  result->n_inliers = st[1];
  result->sumq = st[2];
  result->n_corr = st[3];
  result->fitness = st[3] > 0 ? (double)st[1] / (double)st[3] : 0.0;
  const double tau2 = thresh * thresh;
  result->rmse = st[1] > 0 ? sqrt(((double)st[2] / 1099511627776.0) * tau2 / (double)st[1]) : 0.0;
}

static int register_impl(vfmreg_ctx* ctx, const float* src_xyz, const float* tgt_xyz, const float* src_feats,
                         const float* tgt_feats, int64_t n, int64_t m, int32_t d, const vfmreg_register_params* p,
                         const int32_t* sample_idx, int32_t* corr_dev, uint8_t* mask_dev, bool outputs_on_host,
                         int32_t* corr_host, uint8_t* mask_host, vfmreg_register_result* result) {
  VFM_TRY(arena_reserve(ctx, register_scratch(ctx, n, m, d, p)));
  RegOut out;
  out.corr = corr_dev ? corr_dev : arena_take<int32_t>(ctx, (size_t)n * 2);
  out.mask = mask_dev ? mask_dev : arena_take<uint8_t>(ctx, n);
  out.T = arena_take<double>(ctx, 16);
  out.stats = arena_take<int64_t>(ctx, 8);
  if (!out.corr || !out.mask || !out.T || !out.stats) {
    set_error("register: scratch arena too small");
    return VFMREG_ERR_ALLOC;
  }
  VFM_TRY(register_enqueue(ctx, src_xyz, tgt_xyz, src_feats, tgt_feats, n, m, d, p, sample_idx, out));
  // small results -> pinned host staging -> caller
  VFM_TRY(ensure_pinned(ctx, 256));
  char* pin = static_cast<char*>(ctx->pinned);
  VFM_CUDA(cudaMemcpyAsync(pin, out.T, 16 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  VFM_CUDA(cudaMemcpyAsync(pin + 128, out.stats, 5 * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (outputs_on_host) {
    if (corr_host) VFM_CUDA(cudaMemcpyAsync(corr_host, out.corr, (size_t)n * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (mask_host) VFM_CUDA(cudaMemcpyAsync(mask_host, out.mask, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  }
  VFM_CUDA(cudaStreamSynchronize(ctx->stream));
  fill_result(result, pin, p->inlier_thresh);
  return VFMREG_OK;
}

// VFMREG_MATCH_STREAMS=0 keeps the candidate-search kernels on their lane's stream (A/B comparison)
static bool g_match_streams = [] { const char* e = getenv("VFMREG_MATCH_STREAMS"); return !(e && e[0] == '0'); }();

static int ensure_lanes(vfmreg_ctx* ctx, int lanes) {
  if (!ctx->ev_fork) VFM_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  if (!ctx->match_stream_owned[0]) {
    int lo = 0, hi = 0;
    VFM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // hi = greatest priority (numerically lowest)
    for (int i = 0; i < 2; ++i) VFM_CUDA(cudaStreamCreateWithPriority(&ctx->match_stream_owned[i], cudaStreamNonBlocking, hi));
    for (int i = 0; i < vfmreg_ctx::MATCH_EVENTS; ++i) VFM_CUDA(cudaEventCreateWithFlags(&ctx->match_ev[i], cudaEventDisableTiming));
  }
  for (int l = 1; l < lanes; ++l) {
    if (ctx->lane_stream[l]) continue;
    VFM_CUDA(cudaStreamCreateWithFlags(&ctx->lane_stream[l], cudaStreamNonBlocking));
    VFM_CUDA(cudaEventCreateWithFlags(&ctx->ev_join[l], cudaEventDisableTiming));
  }
  return VFMREG_OK;
}

// Restores the context's stream when a batch entry point leaves (also on an error return in the middle of a batch).
struct StreamGuard {
  vfmreg_ctx* ctx;
  cudaStream_t saved;
  explicit StreamGuard(vfmreg_ctx* c) : ctx(c), saved(c->stream) {}
  ~StreamGuard() {
    ctx->stream = saved;
    ctx->match_stream[0] = ctx->match_stream[1] = nullptr;
  }
};

static int check_register_args(vfmreg_ctx* ctx, const void* a, const void* b, const void* c, const void* e, int64_t n,
                               int64_t m, int32_t d, const vfmreg_register_params* p, vfmreg_register_result* r) {
  VFM_CHECK_ARG(ctx, "null context");
  VFM_CHECK_ARG(a && b && c && e && p && r, "register: null pointer");
  VFM_CHECK_ARG(n > 0 && m > 0 && d > 0, "register: empty input (n=%lld m=%lld d=%d)", (long long)n, (long long)m, d);
  VFM_CHECK_ARG(n < (1LL << 30) && m < (1LL << 30), "register: more than 2^30 points");
  VFM_CHECK_ARG(p->n_hyp > 0, "register: n_hyp must be positive");
  VFM_CHECK_ARG(p->inlier_thresh > 0, "register: inlier_thresh must be > 0");
  return VFMREG_OK;
}

int vfmreg_register(vfmreg_ctx* ctx, const float* src_xyz, const float* tgt_xyz, const float* src_feats,
                    const float* tgt_feats, int64_t n, int64_t m, int32_t d, const vfmreg_register_params* params,
                    const int32_t* sample_idx, int32_t* corr_out, uint8_t* mask_out, vfmreg_register_result* result) {
  VFM_TRY(check_register_args(ctx, src_xyz, tgt_xyz, src_feats, tgt_feats, n, m, d, params, result));
  VFM_CUDA(cudaSetDevice(ctx->device));
  arena_reset(ctx);
  return register_impl(ctx, src_xyz, tgt_xyz, src_feats, tgt_feats, n, m, d, params, sample_idx, corr_out, mask_out, false,
                       nullptr, nullptr, result);
}

int vfmreg_register_host(vfmreg_ctx* ctx, const float* src_xyz, const float* tgt_xyz, const float* src_feats,
                         const float* tgt_feats, int64_t n, int64_t m, int32_t d, const vfmreg_register_params* params,
                         const int32_t* sample_idx, int32_t* corr_out, uint8_t* mask_out,
                         vfmreg_register_result* result) {
  VFM_TRY(check_register_args(ctx, src_xyz, tgt_xyz, src_feats, tgt_feats, n, m, d, params, result));
  VFM_CUDA(cudaSetDevice(ctx->device));
  // device copies of the inputs live in a persistent buffer (grown on demand), separate from the scratch arena
  const size_t b_sx = arena_bytes((size_t)n * 3, 4), b_tx = arena_bytes((size_t)m * 3, 4);
  const size_t b_sf = arena_bytes((size_t)n * d, 4), b_tf = arena_bytes((size_t)m * d, 4);
  const size_t b_si = sample_idx ? arena_bytes((size_t)params->n_hyp * 3, 4) : 0;
  const size_t total = b_sx + b_tx + b_sf + b_tf + b_si;
  if (total > ctx->hbuf_cap) {
    VFM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->hbuf) VFM_CUDA(cudaFree(ctx->hbuf));
    ctx->hbuf = nullptr;
    ctx->hbuf_cap = 0;
    cudaError_t e = cudaMalloc(&ctx->hbuf, total);
    if (e != cudaSuccess) {
      set_error("register_host: cudaMalloc(%zu) failed: %s", total, cudaGetErrorString(e));
      return VFMREG_ERR_ALLOC;
    }
    ctx->hbuf_cap = total;
  }
  char* p = ctx->hbuf;
  float* d_sx = (float*)p; p += b_sx;
  float* d_tx = (float*)p; p += b_tx;
  float* d_sf = (float*)p; p += b_sf;
  float* d_tf = (float*)p; p += b_tf;
  int32_t* d_si = sample_idx ? (int32_t*)p : nullptr;
  VFM_CUDA(cudaMemcpyAsync(d_sx, src_xyz, (size_t)n * 3 * 4, cudaMemcpyHostToDevice, ctx->stream));
  VFM_CUDA(cudaMemcpyAsync(d_tx, tgt_xyz, (size_t)m * 3 * 4, cudaMemcpyHostToDevice, ctx->stream));
  VFM_CUDA(cudaMemcpyAsync(d_sf, src_feats, (size_t)n * d * 4, cudaMemcpyHostToDevice, ctx->stream));
  VFM_CUDA(cudaMemcpyAsync(d_tf, tgt_feats, (size_t)m * d * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (sample_idx)
    VFM_CUDA(cudaMemcpyAsync(d_si, sample_idx, (size_t)params->n_hyp * 3 * 4, cudaMemcpyHostToDevice, ctx->stream));
  arena_reset(ctx);
  return register_impl(ctx, d_sx, d_tx, d_sf, d_tf, n, m, d, params, d_si, nullptr, nullptr, true, corr_out, mask_out,
                       result);
}


int vfmreg_register_batch_host(vfmreg_ctx* ctx, int32_t n_pairs, const float* const* src_xyz, const float* const* tgt_xyz,
                               const float* const* src_feats, const float* const* tgt_feats, const int64_t* n, const int64_t* m,
                               int32_t d, const vfmreg_register_params* params, const int32_t* const* sample_idx,
                               int32_t* const* corr_out, uint8_t* const* mask_out, vfmreg_register_result* results) {
  VFM_CHECK_ARG(ctx && n_pairs > 0 && src_xyz && tgt_xyz && src_feats && tgt_feats && n && m && params && results,
                "register_batch_host: null pointer / empty batch");
  size_t stage_bytes = 0, scratch = 0;
  int64_t n_max = 0;
  for (int i = 0; i < n_pairs; ++i) {
    VFM_TRY(check_register_args(ctx, src_xyz[i], tgt_xyz[i], src_feats[i], tgt_feats[i], n[i], m[i], d, params, results + i));
    const size_t b = arena_bytes((size_t)n[i] * 3, 4) + arena_bytes((size_t)m[i] * 3, 4) + arena_bytes((size_t)n[i] * d, 4) +
                     arena_bytes((size_t)m[i] * d, 4) + (sample_idx ? arena_bytes((size_t)params->n_hyp * 3, 4) : 0);
    stage_bytes = b > stage_bytes ? b : stage_bytes;
    const size_t sc = register_scratch(ctx, n[i], m[i], d, params);
    scratch = sc > scratch ? sc : scratch;
    n_max = n[i] > n_max ? n[i] : n_max;
  }
  VFM_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->copy_stream) {
    VFM_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      VFM_CUDA(cudaEventCreateWithFlags(&ctx->ev_ready[i], cudaEventDisableTiming));
      VFM_CUDA(cudaEventCreateWithFlags(&ctx->ev_consumed[i], cudaEventDisableTiming));
    }
  }
  // persistent device buffer: 2 input stages + 2 (corr, mask) output stages + per-pair (T, stats) slots
  const size_t out_bytes = arena_bytes((size_t)n_max * 2, 4) + arena_bytes(n_max, 1);
  const size_t total = 2 * stage_bytes + 2 * out_bytes + (size_t)n_pairs * 256;
  if (total > ctx->hbuf_cap) {
    VFM_CUDA(cudaStreamSynchronize(ctx->stream));
    VFM_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    if (ctx->hbuf) VFM_CUDA(cudaFree(ctx->hbuf));
    ctx->hbuf = nullptr;
    ctx->hbuf_cap = 0;
    cudaError_t e = cudaMalloc(&ctx->hbuf, total);
    if (e != cudaSuccess) {
      set_error("register_batch_host: cudaMalloc(%zu) failed: %s", total, cudaGetErrorString(e));
      return VFMREG_ERR_ALLOC;
    }
    ctx->hbuf_cap = total;
  }
  arena_reset(ctx);
  VFM_TRY(arena_reserve(ctx, scratch));
  VFM_TRY(ensure_pinned(ctx, (size_t)n_pairs * 256));
  char* slots = ctx->hbuf + 2 * stage_bytes + 2 * out_bytes;

  struct Staged { float *sx, *tx, *sf, *tf; int32_t* si; };
  auto stage_ptrs = [&](int i, int buf) {
    char* p = ctx->hbuf + (size_t)buf * stage_bytes;
    Staged s;
    s.sx = (float*)p; p += arena_bytes((size_t)n[i] * 3, 4);
    s.tx = (float*)p; p += arena_bytes((size_t)m[i] * 3, 4);
    s.sf = (float*)p; p += arena_bytes((size_t)n[i] * d, 4);
    s.tf = (float*)p; p += arena_bytes((size_t)m[i] * d, 4);
    s.si = (sample_idx && sample_idx[i]) ? (int32_t*)p : nullptr;
    return s;
  };
  auto enqueue_h2d = [&](int i) -> int {
    const int buf = i & 1;
    if (i >= 2) VFM_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[buf], 0));  // stage free again
    const Staged s = stage_ptrs(i, buf);
    VFM_CUDA(cudaMemcpyAsync(s.sx, src_xyz[i], (size_t)n[i] * 3 * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    VFM_CUDA(cudaMemcpyAsync(s.tx, tgt_xyz[i], (size_t)m[i] * 3 * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    VFM_CUDA(cudaMemcpyAsync(s.sf, src_feats[i], (size_t)n[i] * d * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    VFM_CUDA(cudaMemcpyAsync(s.tf, tgt_feats[i], (size_t)m[i] * d * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    if (s.si) VFM_CUDA(cudaMemcpyAsync(s.si, sample_idx[i], (size_t)params->n_hyp * 3 * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    VFM_CUDA(cudaEventRecord(ctx->ev_ready[buf], ctx->copy_stream));
    return VFMREG_OK;
  };
  // the previous batch may still be reading the stages: order this batch's first copies after everything enqueued so far
  VFM_CUDA(cudaEventRecord(ctx->ev_consumed[0], ctx->stream));
  VFM_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[0], 0));
  VFM_TRY(enqueue_h2d(0));
  for (int i = 0; i < n_pairs; ++i) {
    const int buf = i & 1;
    if (i + 1 < n_pairs) VFM_TRY(enqueue_h2d(i + 1));  // overlaps with the compute of pair i
    VFM_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_ready[buf], 0));
    const Staged s = stage_ptrs(i, buf);
    RegOut out;
    char* ob = ctx->hbuf + 2 * stage_bytes + (size_t)buf * out_bytes;
    out.corr = (int32_t*)ob;
    out.mask = (uint8_t*)(ob + arena_bytes((size_t)n_max * 2, 4));
    out.T = (double*)(slots + (size_t)i * 256);
    out.stats = (int64_t*)(slots + (size_t)i * 256 + 128);
    arena_reset(ctx);
    VFM_TRY(register_enqueue(ctx, s.sx, s.tx, s.sf, s.tf, n[i], m[i], d, params, s.si, out));
    VFM_CUDA(cudaEventRecord(ctx->ev_consumed[buf], ctx->stream));
    if (corr_out && corr_out[i])
      VFM_CUDA(cudaMemcpyAsync(corr_out[i], out.corr, (size_t)n[i] * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (mask_out && mask_out[i]) VFM_CUDA(cudaMemcpyAsync(mask_out[i], out.mask, (size_t)n[i], cudaMemcpyDeviceToHost, ctx->stream));
  }
  VFM_CUDA(cudaMemcpyAsync(ctx->pinned, slots, (size_t)n_pairs * 256, cudaMemcpyDeviceToHost, ctx->stream));
  VFM_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n_pairs; ++i) fill_result(results + i, static_cast<const char*>(ctx->pinned) + (size_t)i * 256, params->inlier_thresh);
  return VFMREG_OK;
}


int vfmreg_register_batch(vfmreg_ctx* ctx, int32_t n_pairs, const float* const* src_xyz, const float* const* tgt_xyz,
                          const float* const* src_feats, const float* const* tgt_feats, const int64_t* n, const int64_t* m,
                          int32_t d, const vfmreg_register_params* params, const int32_t* const* sample_idx,
                          int32_t* const* corr_out, uint8_t* const* mask_out, vfmreg_register_result* results) {
  VFM_CHECK_ARG(ctx && n_pairs > 0 && src_xyz && tgt_xyz && src_feats && tgt_feats && n && m && params && results,
                "register_batch: null pointer / empty batch");
  size_t scratch = 0;
  int64_t n_max = 0;
  for (int i = 0; i < n_pairs; ++i) {
    VFM_TRY(check_register_args(ctx, src_xyz[i], tgt_xyz[i], src_feats[i], tgt_feats[i], n[i], m[i], d, params, results + i));
    const size_t sc = register_scratch(ctx, n[i], m[i], d, params);
    scratch = sc > scratch ? sc : scratch;
    n_max = n[i] > n_max ? n[i] : n_max;
  }
  VFM_CUDA(cudaSetDevice(ctx->device));
  // scratch arena: [per-pair (T, stats) slots | per lane: fallback corr/mask + per-pair scratch (reused in stream order)]
  const int lanes = ctx->lanes < n_pairs ? ctx->lanes : n_pairs;
  const size_t slots_bytes = (size_t)n_pairs * 256;
  const size_t out_bytes = arena_bytes((size_t)n_max * 2, 4) + arena_bytes(n_max, 1);
  const size_t lane_bytes = out_bytes + scratch + 4096;
  arena_reset(ctx);
  VFM_TRY(arena_reserve(ctx, slots_bytes + lanes * lane_bytes + 4096));
  VFM_TRY(ensure_pinned(ctx, slots_bytes));
  char* slots = arena_take<char>(ctx, slots_bytes);
  if (!slots) {
    set_error("register_batch: scratch arena too small");
    return VFMREG_ERR_ALLOC;
  }
  const size_t mark = ctx->arena.off;
  StreamGuard guard(ctx);
  cudaStream_t lane_streams[vfmreg_ctx::MAX_LANES];
  for (int l = 0; l < vfmreg_ctx::MAX_LANES; ++l) lane_streams[l] = ctx->stream;
  if (lanes > 1) {
    VFM_TRY(ensure_lanes(ctx, lanes));
    VFM_CUDA(cudaEventRecord(ctx->ev_fork, guard.saved));          // inputs are ready in the caller's stream order
    for (int l = 1; l < lanes; ++l) {
      lane_streams[l] = ctx->lane_stream[l];
      VFM_CUDA(cudaStreamWaitEvent(ctx->lane_stream[l], ctx->ev_fork, 0));
    }
    if (g_match_streams) {
      ctx->match_stream[0] = ctx->match_stream_owned[0];
      ctx->match_stream[1] = ctx->match_stream_owned[1];
    }
  }
  for (int i = 0; i < n_pairs; ++i) {
    const int lane = i % lanes;
    ctx->stream = lane_streams[lane];
    ctx->arena.off = mark + (size_t)lane * lane_bytes;   // every pair of a lane reuses the lane's region (stream order)
    int32_t* corr_fb = arena_take<int32_t>(ctx, (size_t)n_max * 2);
    uint8_t* mask_fb = arena_take<uint8_t>(ctx, n_max);
    if (!corr_fb || !mask_fb) {
      set_error("register_batch: scratch arena too small");
      return VFMREG_ERR_ALLOC;
    }
    RegOut out;
    out.corr = (corr_out && corr_out[i]) ? corr_out[i] : corr_fb;
    out.mask = (mask_out && mask_out[i]) ? mask_out[i] : mask_fb;
    out.T = (double*)(slots + (size_t)i * 256);
    out.stats = (int64_t*)(slots + (size_t)i * 256 + 128);
    VFM_TRY(register_enqueue(ctx, src_xyz[i], tgt_xyz[i], src_feats[i], tgt_feats[i], n[i], m[i], d, params,
                             sample_idx ? sample_idx[i] : nullptr, out));
  }
  ctx->stream = guard.saved;
  ctx->match_stream[0] = ctx->match_stream[1] = nullptr;   // every search was handed back to its lane by an event
  for (int l = 1; l < lanes; ++l) {
    VFM_CUDA(cudaEventRecord(ctx->ev_join[l], ctx->lane_stream[l]));
    VFM_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[l], 0));
  }
  VFM_CUDA(cudaMemcpyAsync(ctx->pinned, slots, slots_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  VFM_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n_pairs; ++i) fill_result(results + i, static_cast<const char*>(ctx->pinned) + (size_t)i * 256, params->inlier_thresh);
  return VFMREG_OK;
}

}  // extern "C"
