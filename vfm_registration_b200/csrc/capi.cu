// extern "C" surface of libvfmreg_b200.so (see include/vfmreg_b200.h).
#include <math.h>
#include <stdarg.h>

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

void group_begin(vfmreg_ctx* ctx, int group) {
  if (!ctx->timing) return;
  if (ctx->group_pending[group]) {  // fold the previous interval in before reusing the events
    float ms = 0.f;
    if (cudaEventSynchronize(ctx->ev1[group]) == cudaSuccess &&
        cudaEventElapsedTime(&ms, ctx->ev0[group], ctx->ev1[group]) == cudaSuccess)
      ctx->group_ms[group] += ms;
    ctx->group_pending[group] = 0;
  }
  cudaEventRecord(ctx->ev0[group], ctx->stream);
}

void group_end(vfmreg_ctx* ctx, int group, int n_launches) {
  if (!ctx->timing) return;
  cudaEventRecord(ctx->ev1[group], ctx->stream);
  ctx->group_pending[group] = 1;
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

static bool use_tc(uint32_t flags) {
  const uint32_t algo = flags & VFMREG_ALGO_MASK;
  // the tcgen05 path's error bound assumes unit-norm rows; un-normalised searches stay on the exact fp32 kernel
  return (flags & VFMREG_NORMALIZE) && algo != VFMREG_ALGO_SIMT;
}

static inline int padded_dim(int d, uint32_t flags) { return round_up(d, use_tc(flags) ? 64 : 16); }

// scratch needed by match_nn_impl beyond the caller-visible outputs
static size_t match_scratch(vfmreg_ctx* ctx, int64_t n, int64_t m, int d, uint32_t flags) {
  const int dp = padded_dim(d, flags);
  const bool mutual = (flags & VFMREG_MUTUAL) != 0;
  size_t s = arena_bytes((size_t)n * dp, 4) + arena_bytes((size_t)m * dp, 4);
  if (use_tc(flags)) {
    s += arena_bytes((size_t)n * dp, 2) + arena_bytes((size_t)m * dp, 2) + arena_bytes(n, 1) + arena_bytes(m, 1);
    s += match_tc_scratch(ctx, n, m);
    if (mutual) s += match_tc_scratch(ctx, m, n);
  } else {
    s += match_simt_scratch(ctx, n, m);
    if (mutual) s += match_simt_scratch(ctx, m, n);
  }
  return s;
}

static int match_nn_impl(vfmreg_ctx* ctx, const float* a, int64_t n, const float* b, int64_t m, int32_t d, uint32_t flags,
                         int32_t* idx01, float* sim01, float* sec01, int32_t* idx10, float* sim10, float* sec10) {
  const bool tc = use_tc(flags);
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
  for (int g = 0; g < NUM_GROUPS; ++g) {
    cudaEventCreate(&ctx->ev0[g]);
    cudaEventCreate(&ctx->ev1[g]);
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
  for (int g = 0; g < NUM_GROUPS; ++g) {
    cudaEventDestroy(ctx->ev0[g]);
    cudaEventDestroy(ctx->ev1[g]);
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

int vfmreg_enable_timing(vfmreg_ctx* ctx, int on) {
  VFM_CHECK_ARG(ctx, "null context");
  ctx->timing = on;
  for (int g = 0; g < NUM_GROUPS; ++g) {
    ctx->group_ms[g] = 0.f;
    ctx->group_launches[g] = 0;
    ctx->group_pending[g] = 0;
  }
  return VFMREG_OK;
}

int vfmreg_group_time_ms(vfmreg_ctx* ctx, int group, float* ms_total, int* launches) {
  VFM_CHECK_ARG(ctx && group >= 0 && group < NUM_GROUPS, "bad group");
  if (ctx->group_pending[group]) {
    float ms = 0.f;
    VFM_CUDA(cudaEventSynchronize(ctx->ev1[group]));
    VFM_CUDA(cudaEventElapsedTime(&ms, ctx->ev0[group], ctx->ev1[group]));
    ctx->group_ms[group] += ms;
    ctx->group_pending[group] = 0;
  }
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

static int register_impl(vfmreg_ctx* ctx, const float* src_xyz, const float* tgt_xyz, const float* src_feats,
                         const float* tgt_feats, int64_t n, int64_t m, int32_t d, const vfmreg_register_params* p,
                         const int32_t* sample_idx, int32_t* corr_dev, uint8_t* mask_dev, bool outputs_on_host,
                         int32_t* corr_host, uint8_t* mask_host, vfmreg_register_result* result) {
  const bool mutual = (p->flags & VFMREG_MUTUAL) != 0;
  const bool use_ratio = !(p->ratio != p->ratio);
  const size_t need = match_scratch(ctx, n, m, d, p->flags) + ransac_scratch((int32_t)n, p->n_hyp) +
                      arena_bytes(n, 4) * 3 + arena_bytes(m, 4) + arena_bytes((size_t)n * 2, 4) + arena_bytes(n, 1) +
                      arena_bytes(16, 8) + arena_bytes(8, 8) + 4096;
  VFM_TRY(arena_reserve(ctx, need));
  int32_t* idx01 = arena_take<int32_t>(ctx, n);
  float* sim01 = arena_take<float>(ctx, n);
  float* sec01 = arena_take<float>(ctx, n);
  int32_t* idx10 = mutual ? arena_take<int32_t>(ctx, m) : nullptr;
  int32_t* corr = corr_dev ? corr_dev : arena_take<int32_t>(ctx, (size_t)n * 2);
  uint8_t* mask = mask_dev ? mask_dev : arena_take<uint8_t>(ctx, n);
  double* T = arena_take<double>(ctx, 16);
  int64_t* stats = arena_take<int64_t>(ctx, 8);  // stats[0..3] + count at [4] (int32 view)
  if (!idx01 || !sim01 || !sec01 || !corr || !mask || !T || !stats || (mutual && !idx10)) {
    set_error("register: scratch arena too small");
    return VFMREG_ERR_ALLOC;
  }
  int32_t* count = reinterpret_cast<int32_t*>(stats + 4);
  VFM_TRY(match_nn_impl(ctx, src_feats, n, tgt_feats, m, d, p->flags, idx01, sim01, use_ratio ? sec01 : nullptr, idx10,
                        nullptr, nullptr));
  VFM_TRY(filter_corr(ctx, idx01, sim01, sec01, idx10, n, p->min_cos, p->ratio, mutual, corr, count));
  VFM_TRY(ransac_solve(ctx, src_xyz, tgt_xyz, 0, corr, count, (int32_t)n, sample_idx, p->n_hyp, p->seed, p->inlier_thresh,
                       p->refit, T, nullptr, nullptr, mask, stats));
  // small results -> pinned host staging -> caller
  VFM_TRY(ensure_pinned(ctx, 256));
  char* pin = static_cast<char*>(ctx->pinned);
  VFM_CUDA(cudaMemcpyAsync(pin, T, 16 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  VFM_CUDA(cudaMemcpyAsync(pin + 128, stats, 5 * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (outputs_on_host) {
    if (corr_host) VFM_CUDA(cudaMemcpyAsync(corr_host, corr, (size_t)n * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (mask_host) VFM_CUDA(cudaMemcpyAsync(mask_host, mask, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  }
  VFM_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(result->T, pin, 16 * sizeof(double));
  const int64_t* st = reinterpret_cast<const int64_t*>(pin + 128);
  result->best_hyp = st[0];
  result->n_inliers = st[1];
  result->sumq = st[2];
  result->n_corr = st[3];
  result->fitness = st[3] > 0 ? (double)st[1] / (double)st[3] : 0.0;
  const double tau2 = p->inlier_thresh * p->inlier_thresh;
  result->rmse = st[1] > 0 ? sqrt(((double)st[2] / 1099511627776.0) * tau2 / (double)st[1]) : 0.0;
  return VFMREG_OK;
}

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

}  // extern "C"
