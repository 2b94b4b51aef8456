// extern "C" surface of libvfmreg_b200.so (see include/vfmreg_b200.h).
#include <math.h>
#include <nvtx3/nvToolsExt.h>
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

// A map kept resident on the device for the scans of one scene (registration_node.py:554-590 builds one local map per
// scene and registers 3-5 scans against it): raw float32 coordinates + the renormalised fp32 / fp16 descriptor rows.
struct vfmreg_map {
  vfmreg_ctx* ctx = nullptr;
  int64_t m = 0;
  int32_t d = 0;
  uint32_t flags = 0;       // VFMREG_NORMALIZE | VFMREG_ALGO_* the rows were prepared with
  char* slab = nullptr;     // one allocation: [xyz | prepared rows]
  float* xyz = nullptr;
  vfm::Prepared prep;
};

namespace vfm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void nvtx_push(const char* name) { nvtxRangePushA(name); }
void nvtx_pop() { nvtxRangePop(); }

int arena_reserve(vfmreg_ctx* ctx, size_t bytes) {
  ctx->arena.limit = 0;
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

static int ensure_hbuf(vfmreg_ctx* ctx, size_t bytes, const char* who) {
  if (bytes <= ctx->hbuf_cap) return VFMREG_OK;
  VFM_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->copy_stream) VFM_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  if (ctx->hbuf) VFM_CUDA(cudaFree(ctx->hbuf));
  ctx->hbuf = nullptr;
  ctx->hbuf_cap = 0;
  cudaError_t e = cudaMalloc(&ctx->hbuf, bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaMalloc(%zu) failed: %s", who, bytes, cudaGetErrorString(e));
    return VFMREG_ERR_ALLOC;
  }
  ctx->hbuf_cap = bytes;
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

// ---- prepared descriptor sets ----------------------------------------------------------------------------------------
// [fp32 rows (rows x dp) | fp16 rows (tensor path) | non-zero flags (tensor path)]
static size_t prepared_bytes(int64_t rows, int d, uint32_t flags) {
  const int dp = padded_dim(d, flags);
  size_t s = arena_bytes((size_t)rows * dp, 4);
  if (use_tc(flags, d)) s += arena_bytes((size_t)rows * dp, 2) + arena_bytes(rows, 1);
  return s;
}

// renormalise (or copy) x (rows x d) into `mem` on the context's stream
static int prepare_into(vfmreg_ctx* ctx, const float* x, int64_t rows, int d, uint32_t flags, char* mem, Prepared* out) {
  const bool tc = use_tc(flags, d);
  const int dp = padded_dim(d, flags);
  float* f32 = reinterpret_cast<float*>(mem);
  uint16_t* f16 = nullptr;
  uint8_t* nz = nullptr;
  if (tc) {
    f16 = reinterpret_cast<uint16_t*>(mem + arena_bytes((size_t)rows * dp, 4));
    nz = reinterpret_cast<uint8_t*>(mem + arena_bytes((size_t)rows * dp, 4) + arena_bytes((size_t)rows * dp, 2));
  }
  NvtxRange r("normalize_rows");
  GroupScope g(ctx, GROUP_NORMALIZE, 1);
  VFM_TRY(normalize_rows(ctx, x, rows, d, dp, (flags & VFMREG_NORMALIZE) != 0, f32, f16, nz));
  out->f32 = f32;
  out->f16 = f16;
  out->nz = nz;
  out->rows = rows;
  out->dp = dp;
  out->tc = tc;
  return VFMREG_OK;
}

static int prepare_arena(vfmreg_ctx* ctx, const float* x, int64_t rows, int d, uint32_t flags, Prepared* out) {
  char* mem = arena_take<char>(ctx, prepared_bytes(rows, d, flags));
  if (!mem) {
    set_error("scratch arena too small for %lld x %d descriptor rows", (long long)rows, d);
    return VFMREG_ERR_ALLOC;
  }
  return prepare_into(ctx, x, rows, d, flags, mem, out);
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

// scratch of the searches of one (a, b) pair beyond the prepared rows and the caller-visible outputs
static size_t search_scratch(vfmreg_ctx* ctx, int64_t n, int64_t m, int d, uint32_t flags, bool pruned) {
  const bool mutual = (flags & VFMREG_MUTUAL) != 0;
  size_t s = 0;
  if (use_tc(flags, d)) {
    s += match_tc_scratch(ctx, n, m);
    if (pruned) s += pruned_scratch(ctx, n, d, flags);
    else if (mutual) s += match_tc_scratch(ctx, m, n);
  } else {
    s += match_simt_scratch(ctx, n, m);
    if (mutual) s += match_simt_scratch(ctx, m, n);
  }
  return s;
}

// both searches with all six outputs (the public match_nn)
static int search_full(vfmreg_ctx* ctx, const Prepared& A, const Prepared& B, uint32_t flags, int32_t* idx01, float* sim01,
                       float* sec01, int32_t* idx10, float* sim10, float* sec10, float floor = NAN) {
  if (flags & VFMREG_MUTUAL) VFM_CHECK_ARG(idx10, "match_nn: VFMREG_MUTUAL needs idx10");
  if (A.tc) {
    VFM_TRY(match_tc(ctx, A.f32, A.f16, A.nz, A.rows, B.f32, B.f16, B.rows, A.dp, idx01, sim01, sec01, nullptr, nullptr, floor));
    if (flags & VFMREG_MUTUAL)
      VFM_TRY(match_tc(ctx, B.f32, B.f16, B.nz, B.rows, A.f32, A.f16, A.rows, A.dp, idx10, sim10, sec10));
  } else {
    VFM_TRY(match_simt(ctx, A.f32, A.rows, B.f32, B.rows, A.dp, idx01, sim01, sec01));
    if (flags & VFMREG_MUTUAL) VFM_TRY(match_simt(ctx, B.f32, B.rows, A.f32, A.rows, A.dp, idx10, sim10, sec10));
  }
  return VFMREG_OK;
}

// register(): forward search -> gate -> (mutual check) -> ordered correspondence list on the device, in three stages so
// that a batch can run a stage for a whole group of pairs before the next one (batch_run):
//   1  renormalise the scan, start the forward search (on the tensor-core path only the search kernel is enqueued);
//   2  re-rank, gate; pruned mutual check: gather the map rows the gated queries point at and start the reverse search
//      over them; otherwise the final correspondence list;
//   3  pruned mutual check: re-rank of the reverse search, keep the candidates whose map row points back.
// A caller that gates on the cosine and does not need the runner-up hands the gate to the candidate search as a recording
// floor: queries that cannot reach it report "no match" instead of their (rejected anyway) best.
struct ScanState {
  Prepared A;
  bool pruned = false, full_reverse = false;
  int32_t *idx01 = nullptr, *idx10 = nullptr;
  float *sim01 = nullptr, *sec01 = nullptr;
  int32_t *cand = nullptr, *cand_count = nullptr, *back = nullptr;
  float *sel32 = nullptr, *selsim = nullptr;
  uint16_t* sel16 = nullptr;
  uint8_t* selnz = nullptr;
  TcPending fwd, rev;
  size_t arena_off = 0, arena_limit = 0;   // where the next stage goes on carving the pair's scratch region
};

static inline void state_save(vfmreg_ctx* ctx, ScanState* st) {
  st->arena_off = ctx->arena.off;
  st->arena_limit = ctx->arena.limit;
}
static inline void state_restore(vfmreg_ctx* ctx, const ScanState& st) {
  ctx->arena.off = st.arena_off;
  ctx->arena.limit = st.arena_limit;
}

static int search_stage1(vfmreg_ctx* ctx, const Prepared& A, const Prepared& B, const vfmreg_register_params* p, ScanState* st) {
  const int64_t n = A.rows, m = B.rows;
  const bool mutual = (p->flags & VFMREG_MUTUAL) != 0;
  const bool use_ratio = !(p->ratio != p->ratio);
  const float floor = use_ratio ? NAN : p->min_cos;   // NAN when there is no gate either
  st->A = A;
  st->pruned = A.tc && mutual && n <= m && !g_full_mutual;
  st->full_reverse = mutual && !st->pruned;
  st->idx01 = arena_take<int32_t>(ctx, n);
  st->sim01 = arena_take<float>(ctx, n);
  st->sec01 = use_ratio ? arena_take<float>(ctx, n) : nullptr;
  st->idx10 = st->full_reverse ? arena_take<int32_t>(ctx, m) : nullptr;
  if (!st->idx01 || !st->sim01 || (use_ratio && !st->sec01) || (st->full_reverse && !st->idx10)) {
    set_error("register: scratch arena too small");
    return VFMREG_ERR_ALLOC;
  }
  if (st->pruned) {
    const int dp = A.dp;
    st->cand = arena_take<int32_t>(ctx, (size_t)n * 2);
    st->cand_count = arena_take<int32_t>(ctx, 1);
    st->sel32 = arena_take<float>(ctx, (size_t)n * dp);
    st->sel16 = arena_take<uint16_t>(ctx, (size_t)n * dp);
    st->selnz = arena_take<uint8_t>(ctx, n);
    st->back = arena_take<int32_t>(ctx, n);
    st->selsim = arena_take<float>(ctx, n);
    if (!st->cand || !st->cand_count || !st->sel32 || !st->sel16 || !st->selnz || !st->back || !st->selsim) {
      set_error("register: scratch arena too small");
      return VFMREG_ERR_ALLOC;
    }
  }
  if (A.tc) {
    VFM_TRY(match_tc_begin(ctx, A.f32, A.f16, A.nz, n, B.f32, B.f16, m, A.dp, st->idx01, st->sim01, st->sec01, nullptr, nullptr, floor,
                           &st->fwd));
    if (st->full_reverse)
      VFM_TRY(match_tc_begin(ctx, B.f32, B.f16, B.nz, m, A.f32, A.f16, n, A.dp, st->idx10, nullptr, nullptr, nullptr, nullptr, NAN,
                             &st->rev));
  } else {
    VFM_TRY(match_simt(ctx, A.f32, n, B.f32, m, A.dp, st->idx01, st->sim01, st->sec01));
    if (st->full_reverse) VFM_TRY(match_simt(ctx, B.f32, m, A.f32, n, A.dp, st->idx10, nullptr, nullptr));
  }
  state_save(ctx, st);
  return VFMREG_OK;
}

// `searches_done`: the event behind the last search of the group (null: wait for this pair's own searches)
static int search_stage2(vfmreg_ctx* ctx, const Prepared& B, const vfmreg_register_params* p, int32_t* corr, int32_t* count,
                         ScanState* st, cudaEvent_t searches_done) {
  state_restore(ctx, *st);
  const Prepared& A = st->A;
  const int64_t n = A.rows;
  if (A.tc) {
    VFM_TRY(match_tc_finish(ctx, st->fwd, searches_done));
    if (st->full_reverse) VFM_TRY(match_tc_finish(ctx, st->rev, searches_done));
  }
  if (st->pruned) {
    NvtxRange r("gate_and_gather");
    {
      GroupScope g(ctx, GROUP_FILTER, 1);
      VFM_TRY(filter_corr(ctx, st->idx01, st->sim01, st->sec01, nullptr, n, p->min_cos, p->ratio, 0, st->cand, st->cand_count));
    }
    {
      GroupScope g(ctx, GROUP_GATHER, 1);
      // <b_j, a_i> has the same canonical value as <a_i, b_j> = sim01[i]: a lower bound of row j's best that starts the
      // candidate recording of the reverse search near the answer
      VFM_TRY(gather_rows(ctx, st->cand, st->cand_count, n, 1, A.dp, B.f32, B.f16, B.nz, st->sel32, st->sel16, st->selnz, st->sim01,
                          st->selsim));
    }
    VFM_TRY(match_tc_begin(ctx, st->sel32, st->sel16, st->selnz, n, A.f32, A.f16, n, A.dp, st->back, nullptr, nullptr, st->cand_count,
                           st->selsim, NAN, &st->rev));
  } else {
    NvtxRange r("filter_corr");
    GroupScope g(ctx, GROUP_FILTER, 1);
    VFM_TRY(filter_corr(ctx, st->idx01, st->sim01, st->sec01, st->idx10, n, p->min_cos, p->ratio, (p->flags & VFMREG_MUTUAL) != 0, corr,
                        count));
  }
  state_save(ctx, st);
  return VFMREG_OK;
}

static int search_stage3(vfmreg_ctx* ctx, int32_t* corr, int32_t* count, ScanState* st, cudaEvent_t searches_done) {
  state_restore(ctx, *st);
  if (!st->pruned) return VFMREG_OK;
  VFM_TRY(match_tc_finish(ctx, st->rev, searches_done));
  NvtxRange r("mutual_check");
  GroupScope g(ctx, GROUP_MUTUAL, 1);
  VFM_TRY(filter_mutual_list(ctx, st->cand, st->cand_count, st->back, st->A.rows, corr, count));
  state_save(ctx, st);
  return VFMREG_OK;
}

}  // namespace vfm

using namespace vfm;

struct RegOut {
  double* T;        // device, 16
  int64_t* stats;   // device, 8 (stats[0..3] + correspondence count as int32 at [4])
  int32_t* corr;    // device, n x 2
  uint8_t* mask;    // device, n
};

// scratch of one scan against an already prepared map
static size_t scan_scratch(vfmreg_ctx* ctx, int64_t n, int64_t m, int32_t d, const vfmreg_register_params* p) {
  const bool mutual = (p->flags & VFMREG_MUTUAL) != 0;
  const bool pruned = prune_mutual(n, m, d, p->flags);
  return prepared_bytes(n, d, p->flags) + search_scratch(ctx, n, m, d, p->flags, pruned) + arena_bytes(n, 4) * 3 +
         ((mutual && !pruned) ? arena_bytes(m, 4) : 0) + ransac_scratch((int32_t)n, p->n_hyp) + arena_bytes((size_t)n * 2, 4) +
         arena_bytes(n, 1) + arena_bytes(16, 8) + arena_bytes(8, 8) + 4096;
}

// One scan against a prepared map, enqueued on ctx->stream stage by stage (no host synchronisation; the arena must
// already be reserved); stage 3 ends with the RANSAC solve.
static int scan_stage1(vfmreg_ctx* ctx, const float* src_feats, int64_t n, int32_t d, const Prepared& B, const vfmreg_register_params* p,
                       ScanState* st) {
  Prepared A;
  VFM_TRY(prepare_arena(ctx, src_feats, n, d, p->flags, &A));
  return search_stage1(ctx, A, B, p, st);
}

static int scan_stage3(vfmreg_ctx* ctx, const float* src_xyz, int64_t n, const float* tgt_xyz, const vfmreg_register_params* p,
                       const int32_t* sample_idx, const RegOut& out, ScanState* st, cudaEvent_t searches_done) {
  int32_t* count = reinterpret_cast<int32_t*>(out.stats + 4);
  VFM_TRY(search_stage3(ctx, out.corr, count, st, searches_done));
  NvtxRange r("ransac");
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

// VFMREG_MATCH_STREAMS=0 keeps every candidate search on its lane's stream instead of the batch's search stream (A/B)
static bool g_match_streams = [] { const char* e = getenv("VFMREG_MATCH_STREAMS"); return !(e && e[0] == '0'); }();

static int ensure_lanes(vfmreg_ctx* ctx, int lanes) {
  if (!ctx->ev_fork) {
    VFM_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    VFM_CUDA(cudaEventCreateWithFlags(&ctx->ev_join[0], cudaEventDisableTiming));
  }
  if (!ctx->match_stream_owned) {
    int lo = 0, hi = 0;
    VFM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // hi = greatest priority (numerically lowest)
    VFM_CUDA(cudaStreamCreateWithPriority(&ctx->match_stream_owned, cudaStreamNonBlocking, hi));
    for (int i = 0; i < vfmreg_ctx::MATCH_EVENTS; ++i) VFM_CUDA(cudaEventCreateWithFlags(&ctx->match_ev[i], cudaEventDisableTiming));
  }
  for (int l = 1; l < lanes; ++l) {
    if (ctx->lane_stream[l]) continue;
    VFM_CUDA(cudaStreamCreateWithFlags(&ctx->lane_stream[l], cudaStreamNonBlocking));
    VFM_CUDA(cudaEventCreateWithFlags(&ctx->ev_join[l], cudaEventDisableTiming));
  }
  return VFMREG_OK;
}

static int ensure_map_events(vfmreg_ctx* ctx, int count) {
  if (count <= ctx->map_ev_cap) return VFMREG_OK;
  cudaEvent_t* ev = static_cast<cudaEvent_t*>(realloc(ctx->map_ev, sizeof(cudaEvent_t) * count));
  if (!ev) {
    set_error("out of host memory");
    return VFMREG_ERR_ALLOC;
  }
  ctx->map_ev = ev;
  for (int i = ctx->map_ev_cap; i < count; ++i) {
    ctx->map_ev[i] = nullptr;
    VFM_CUDA(cudaEventCreateWithFlags(&ctx->map_ev[i], cudaEventDisableTiming));
    ctx->map_ev_cap = i + 1;
  }
  return VFMREG_OK;
}

static int ensure_copy_stream(vfmreg_ctx* ctx) {
  if (ctx->copy_stream) return VFMREG_OK;
  VFM_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    VFM_CUDA(cudaEventCreateWithFlags(&ctx->ev_ready[i], cudaEventDisableTiming));
    VFM_CUDA(cudaEventCreateWithFlags(&ctx->ev_consumed[i], cudaEventDisableTiming));
    VFM_CUDA(cudaEventCreateWithFlags(&ctx->ev_target_free[i], cudaEventDisableTiming));
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
    ctx->match_stream = nullptr;
    ctx->arena.limit = 0;
  }
};

static int check_register_args(vfmreg_ctx* ctx, const void* a, const void* b, const void* c, const void* e, int64_t n,
                               int64_t m, int32_t d, const vfmreg_register_params* p, const void* r) {
  VFM_CHECK_ARG(ctx, "null context");
  VFM_CHECK_ARG(a && b && c && e && p && r, "register: null pointer");
  VFM_CHECK_ARG(n > 0 && m > 0 && d > 0, "register: empty input (n=%lld m=%lld d=%d)", (long long)n, (long long)m, d);
  VFM_CHECK_ARG(n < (1LL << 30) && m < (1LL << 30), "register: more than 2^30 points");
  VFM_CHECK_ARG(p->n_hyp > 0, "register: n_hyp must be positive");
  VFM_CHECK_ARG(p->inlier_thresh > 0, "register: inlier_thresh must be > 0");
  return VFMREG_OK;
}

// One entry of a batch: a scan, and the map it is registered against -- either raw arrays (prepared inside the call; a
// run of consecutive pairs with the same tgt_feats / tgt_xyz / m shares one preparation) or a resident map.
struct BatchView {
  int32_t n_pairs;
  const float* const* src_xyz;
  const float* const* src_feats;
  const int64_t* n;
  const float* const* tgt_xyz;     // null when `map` is set
  const float* const* tgt_feats;
  const int64_t* m;
  const vfmreg_map* map;           // resident map shared by every pair, or null
  int32_t d;
  const int32_t* const* sample_idx;
  int32_t* const* corr_out;
  uint8_t* const* mask_out;
  int64_t m_of(int i) const { return map ? map->m : m[i]; }
  bool same_target(int i, int j) const {
    return map || (tgt_feats[i] == tgt_feats[j] && tgt_xyz[i] == tgt_xyz[j] && m[i] == m[j]);
  }
};

static int check_batch(vfmreg_ctx* ctx, const BatchView& b, const vfmreg_register_params* p, const vfmreg_register_result* results,
                       size_t* scratch, int64_t* n_max, size_t* target_bytes, int* n_runs) {
  VFM_CHECK_ARG(ctx && b.n_pairs > 0 && b.src_xyz && b.src_feats && b.n && p && results, "register batch: null pointer / empty batch");
  VFM_CHECK_ARG(b.map || (b.tgt_xyz && b.tgt_feats && b.m), "register batch: null target arrays");
  if (b.map) VFM_CHECK_ARG(b.map->ctx == ctx && b.map->d == b.d, "register_scans: the map belongs to another context or has %d-d descriptors", b.map->d);
  *scratch = 0;
  *n_max = 0;
  *target_bytes = 0;
  *n_runs = 0;
  for (int i = 0; i < b.n_pairs; ++i) {
    const int64_t m = b.m_of(i);
    VFM_TRY(check_register_args(ctx, b.src_xyz[i], b.map ? (const void*)b.map : (const void*)b.tgt_xyz[i], b.src_feats[i],
                                b.map ? (const void*)b.map : (const void*)b.tgt_feats[i], b.n[i], m, b.d, p, results));
    const size_t sc = scan_scratch(ctx, b.n[i], m, b.d, p);
    *scratch = sc > *scratch ? sc : *scratch;
    *n_max = b.n[i] > *n_max ? b.n[i] : *n_max;
    if (!b.map) {
      const size_t tb = prepared_bytes(m, b.d, p->flags);
      *target_bytes = tb > *target_bytes ? tb : *target_bytes;
      if (i == 0 || !b.same_target(i, i - 1)) *n_runs += 1;
    }
  }
  return VFMREG_OK;
}

// ---- the batch engine --------------------------------------------------------------------------------------------------
// Every pair goes through three stages on its own CUDA stream ("lane"; pair i uses lane i % lanes, lane 0 = the caller's
// stream): 1 = renormalise the scan (and prepare the run's map when the pair opens one) and start the forward search,
// 2 = re-rank, gate, gather and start the pruned reverse search, 3 = re-rank of the reverse search, mutual filter, RANSAC.
// The search kernels of all lanes go to ONE high-priority search stream (match_tc_begin); a search needs every SM (one
// 209 KB CTA each) and more than half of the L2 -> SM bandwidth, so what runs beside it matters -- the two schedules below
// differ in exactly that (measurements: profiles/r2_stage_times.txt, tools/corun.py).
// Prepared maps live in a small ring of slots; a run of consecutive pairs with the same target shares one slot.
//   device inputs: the arrays are used in place.
//   host inputs:   a copy stream uploads the scans of later pairs (and the map of a run, once) while the current ones are
//                  processed: 2 x lanes scan stages (xyz + descriptors + sample indices) and (corr, mask) output stages, raw map
//                  stages beside the prepared-map slots -- persistent device staging (ctx->hbuf), grown on demand.
static bool g_grouped = [] { const char* e = getenv("VFMREG_SCHEDULE"); return e && e[0] == 'g'; }();

static int batch_run(vfmreg_ctx* ctx, const BatchView& b, const vfmreg_register_params* params, vfmreg_register_result* results,
                     bool host) {
  size_t scratch = 0, target_bytes = 0;
  int64_t n_max = 0;
  int n_runs = 0;
  VFM_TRY(check_batch(ctx, b, params, results, &scratch, &n_max, &target_bytes, &n_runs));
  VFM_CUDA(cudaSetDevice(ctx->device));
  const int n_pairs = b.n_pairs, d = b.d;
  const int lanes = ctx->lanes < n_pairs ? ctx->lanes : n_pairs;
  constexpr int MAX_RING = 2 * vfmreg_ctx::MAX_LANES + 2;
  const int ring_cap = host ? 2 * lanes : lanes + 2;
  const int ring = b.map ? 0 : (n_runs < ring_cap ? n_runs : ring_cap);     // prepared-map slots
  const int stages = host ? (n_pairs < 2 * lanes ? n_pairs : 2 * lanes) : 0;  // host: scan / output stages
  size_t scan_stage = 0, raw_stage = 0, xyz_stage = 0;
  if (host) {
    for (int i = 0; i < n_pairs; ++i) {
      const size_t s = arena_bytes((size_t)b.n[i] * 3, 4) + arena_bytes((size_t)b.n[i] * d, 4) +
                       (b.sample_idx ? arena_bytes((size_t)params->n_hyp * 3, 4) : 0);
      scan_stage = s > scan_stage ? s : scan_stage;
      if (!b.map) {
        const size_t r = arena_bytes((size_t)b.m[i] * d, 4), x = arena_bytes((size_t)b.m[i] * 3, 4);
        raw_stage = r > raw_stage ? r : raw_stage;
        xyz_stage = x > xyz_stage ? x : xyz_stage;
      }
    }
  }
  const size_t slots_bytes = (size_t)n_pairs * 256;
  const size_t out_bytes = arena_bytes((size_t)n_max * 2, 4) + arena_bytes(n_max, 1);
  // host staging: [ring x (raw map | map xyz | prepared map) | stages x scan | stages x (corr, mask)]
  const size_t hslot = raw_stage + xyz_stage + target_bytes;
  const size_t o_scan = (size_t)ring * hslot, o_out = o_scan + (size_t)stages * scan_stage;
  if (host) {
    VFM_TRY(ensure_hbuf(ctx, o_out + (size_t)stages * out_bytes, "register_batch_host"));
    VFM_TRY(ensure_copy_stream(ctx));
  }
  // scratch arena: [per-pair (T, stats) slots | device inputs: prepared-map ring | per lane: fallback corr/mask + per-scan scratch]
  const size_t lane_bytes = out_bytes + scratch + 4096;
  arena_reset(ctx);
  VFM_TRY(arena_reserve(ctx, slots_bytes + (host ? 0 : (size_t)ring * target_bytes) + lanes * lane_bytes + 8192));
  VFM_TRY(ensure_pinned(ctx, slots_bytes));
  char* slots = arena_take<char>(ctx, slots_bytes);
  char* ring_mem = (ring && !host) ? arena_take<char>(ctx, (size_t)ring * target_bytes) : nullptr;
  if (!slots || (ring && !host && !ring_mem)) {
    set_error("register_batch: scratch arena too small");
    return VFMREG_ERR_ALLOC;
  }
  const size_t mark = ctx->arena.off;
  StreamGuard guard(ctx);
  cudaStream_t lane_streams[vfmreg_ctx::MAX_LANES];
  for (int l = 0; l < vfmreg_ctx::MAX_LANES; ++l) lane_streams[l] = ctx->stream;
  cudaStream_t ks = nullptr;   // the search stream
  if (lanes > 1) {
    VFM_TRY(ensure_lanes(ctx, lanes));
    VFM_CUDA(cudaEventRecord(ctx->ev_fork, guard.saved));          // inputs are ready in the caller's stream order
    for (int l = 1; l < lanes; ++l) {
      lane_streams[l] = ctx->lane_stream[l];
      VFM_CUDA(cudaStreamWaitEvent(ctx->lane_stream[l], ctx->ev_fork, 0));
    }
    if (g_match_streams) {
      ks = ctx->match_stream_owned;
      VFM_CUDA(cudaStreamWaitEvent(ks, ctx->ev_fork, 0));
      ctx->match_stream = ks;
    }
  }
  // events: ring slot s: [s * EV_PER_SLOT] = "prepared", [.. + 1 + l] = "lane l is done with the slot's current map";
  //         scan stage k (host): [stage_ev + 2k] = "uploaded", [stage_ev + 2k + 1] = "consumed, outputs copied back"
  constexpr int EV_PER_SLOT = 1 + vfmreg_ctx::MAX_LANES;
  const int stage_ev = ring * EV_PER_SLOT;
  VFM_TRY(ensure_map_events(ctx, stage_ev + 2 * stages));
  uint32_t slot_lanes[MAX_RING] = {};   // lanes that used the slot's current map
  Prepared slot_prep[MAX_RING];

  struct Staged { float *sx, *sf; int32_t* si; };
  auto scan_ptrs = [&](int i) {
    char* p = ctx->hbuf + o_scan + (size_t)(i % stages) * scan_stage;
    Staged s;
    s.sx = (float*)p; p += arena_bytes((size_t)b.n[i] * 3, 4);
    s.sf = (float*)p; p += arena_bytes((size_t)b.n[i] * d, 4);
    s.si = (b.sample_idx && b.sample_idx[i]) ? (int32_t*)p : nullptr;
    return s;
  };
  auto is_first = [&](int i) { return !b.map && (i == 0 || !b.same_target(i, i - 1)); };
  // host inputs: upload of pair j on the copy stream (its map first when it opens a run)
  int copy_next = 0, copy_run = -1;
  auto enqueue_h2d = [&](int j) -> int {
    cudaStream_t cs = ctx->copy_stream;
    if (is_first(j)) {
      ++copy_run;
      const int slot = copy_run % ring;
      cudaEvent_t* ev = ctx->map_ev + slot * EV_PER_SLOT;
      for (int l = 0; l < lanes; ++l)      // the slot's previous map may still be in use
        if ((slot_lanes[slot] >> l) & 1u) VFM_CUDA(cudaStreamWaitEvent(cs, ev[1 + l], 0));
      slot_lanes[slot] = 0;
      char* mem = ctx->hbuf + (size_t)slot * hslot;
      VFM_TRY(h2d_copy(ctx, mem, b.tgt_feats[j], (size_t)b.m[j] * d * 4, cs));
      VFM_TRY(h2d_copy(ctx, mem + raw_stage, b.tgt_xyz[j], (size_t)b.m[j] * 3 * 4, cs));
    }
    cudaEvent_t* sev = ctx->map_ev + stage_ev + 2 * (j % stages);
    if (j >= stages) VFM_CUDA(cudaStreamWaitEvent(cs, sev[1], 0));   // stage free again (pair j - stages is done with it)
    const Staged s = scan_ptrs(j);
    VFM_TRY(h2d_copy(ctx, s.sx, b.src_xyz[j], (size_t)b.n[j] * 3 * 4, cs));
    VFM_TRY(h2d_copy(ctx, s.sf, b.src_feats[j], (size_t)b.n[j] * d * 4, cs));
    if (s.si) VFM_TRY(h2d_copy(ctx, s.si, b.sample_idx[j], (size_t)params->n_hyp * 3 * 4, cs));
    VFM_CUDA(cudaEventRecord(sev[0], cs));
    return VFMREG_OK;
  };
  if (host) {
    // the previous call may still be reading the staging buffers: order this batch's copies after everything enqueued so far
    VFM_CUDA(cudaEventRecord(ctx->ev_consumed[0], guard.saved));
    VFM_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[0], 0));
  }
  auto next_event = [&]() {
    cudaEvent_t ev = ctx->match_ev[ctx->match_ev_head];
    ctx->match_ev_head = (ctx->match_ev_head + 1) % vfmreg_ctx::MATCH_EVENTS;
    return ev;
  };
  struct PairCtx {
    Staged st;
    const float* tgt_xyz;
    const Prepared* B;
    int slot;
    RegOut out;
    ScanState state;
  };
  PairCtx pc[vfmreg_ctx::MAX_LANES];   // pair i lives in pc[i % lanes] from its stage 1 to its stage 3
  int run = -1;
  auto stage1 = [&](int i) -> int {
    const int k = i % lanes;
    PairCtx& c = pc[k];
    ctx->stream = lane_streams[k];
    c.st = host ? scan_ptrs(i) : Staged{nullptr, nullptr, nullptr};
    if (host) VFM_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->map_ev[stage_ev + 2 * (i % stages)], 0));
    c.B = b.map ? &b.map->prep : nullptr;
    c.tgt_xyz = b.map ? b.map->xyz : (host ? nullptr : b.tgt_xyz[i]);
    c.slot = -1;
    if (!b.map) {
      const bool first = is_first(i);
      if (first) ++run;
      const int slot = run % ring;
      c.slot = slot;
      cudaEvent_t* ev = ctx->map_ev + slot * EV_PER_SLOT;
      char* hmem = host ? ctx->hbuf + (size_t)slot * hslot : nullptr;
      if (first) {
        if (!host) {   // (host inputs: the copy stream waited for the slot's previous users before overwriting it)
          for (int l = 0; l < lanes; ++l)
            if ((slot_lanes[slot] >> l) & 1u) VFM_CUDA(cudaStreamWaitEvent(ctx->stream, ev[1 + l], 0));
          slot_lanes[slot] = 0;
        }
        VFM_TRY(prepare_into(ctx, host ? (const float*)hmem : b.tgt_feats[i], b.m[i], d, params->flags,
                             host ? hmem + raw_stage + xyz_stage : ring_mem + (size_t)slot * target_bytes, &slot_prep[slot]));
        VFM_CUDA(cudaEventRecord(ev[0], ctx->stream));
      } else {
        VFM_CUDA(cudaStreamWaitEvent(ctx->stream, ev[0], 0));
      }
      c.B = &slot_prep[slot];
      if (host) c.tgt_xyz = (const float*)(hmem + raw_stage);
    }
    // every pair of a lane reuses the lane's region (stream order); a pair may not run past its lane's end
    ctx->arena.off = mark + (size_t)k * lane_bytes;
    ctx->arena.limit = ctx->arena.off + lane_bytes;
    if (host) {
      char* ob = ctx->hbuf + o_out + (size_t)(i % stages) * out_bytes;
      c.out.corr = (int32_t*)ob;
      c.out.mask = (uint8_t*)(ob + arena_bytes((size_t)n_max * 2, 4));
    } else {
      int32_t* corr_fb = arena_take<int32_t>(ctx, (size_t)n_max * 2);
      uint8_t* mask_fb = arena_take<uint8_t>(ctx, n_max);
      if (!corr_fb || !mask_fb) {
        set_error("register_batch: scratch arena too small");
        return VFMREG_ERR_ALLOC;
      }
      c.out.corr = (b.corr_out && b.corr_out[i]) ? b.corr_out[i] : corr_fb;
      c.out.mask = (b.mask_out && b.mask_out[i]) ? b.mask_out[i] : mask_fb;
    }
    c.out.T = (double*)(slots + (size_t)i * 256);
    c.out.stats = (int64_t*)(slots + (size_t)i * 256 + 128);
    return scan_stage1(ctx, host ? c.st.sf : b.src_feats[i], b.n[i], d, *c.B, params, &c.state);
  };
  auto stage2 = [&](int i, cudaEvent_t searches_done) -> int {
    PairCtx& c = pc[i % lanes];
    ctx->stream = lane_streams[i % lanes];
    return search_stage2(ctx, *c.B, params, c.out.corr, reinterpret_cast<int32_t*>(c.out.stats + 4), &c.state, searches_done);
  };
  auto stage3 = [&](int i, cudaEvent_t searches_done) -> int {
    const int k = i % lanes;
    PairCtx& c = pc[k];
    ctx->stream = lane_streams[k];
    VFM_TRY(scan_stage3(ctx, host ? c.st.sx : b.src_xyz[i], b.n[i], c.tgt_xyz, params,
                        host ? c.st.si : (b.sample_idx ? b.sample_idx[i] : nullptr), c.out, &c.state, searches_done));
    if (host) {
      if (b.corr_out && b.corr_out[i])
        VFM_CUDA(cudaMemcpyAsync(b.corr_out[i], c.out.corr, (size_t)b.n[i] * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
      if (b.mask_out && b.mask_out[i])
        VFM_CUDA(cudaMemcpyAsync(b.mask_out[i], c.out.mask, (size_t)b.n[i], cudaMemcpyDeviceToHost, ctx->stream));
      VFM_CUDA(cudaEventRecord(ctx->map_ev[stage_ev + 2 * (i % stages) + 1], ctx->stream));
    }
    if (c.slot >= 0) {
      VFM_CUDA(cudaEventRecord(ctx->map_ev[c.slot * EV_PER_SLOT + 1 + k], ctx->stream));
      slot_lanes[c.slot] |= 1u << k;
    }
    if (lanes > 1) VFM_CUDA(cudaEventRecord(ctx->ev_join[k], ctx->stream));
    return VFMREG_OK;
  };
  if (!g_grouped && lanes >= 3) {
    // PIPELINED schedule (default): stage 1 of pair i, stage 2 of pair i - 1, stage 3 of pair i - 2, ...  The search stream
    // sees F(i), P(i-1), F(i+1), P(i), ... back to back; the small kernels of a pair run beside the searches of its successors
    // (about half of their time is hidden that way, tools/corun.py) and take about 10 % off the search kernel's rate.
    for (int i = 0; i < n_pairs + 2; ++i) {
      if (host)   // pair j reuses the stage of pair j - stages, whose stage 3 must already be enqueued (pairs <= i - 3 are)
        for (; copy_next < n_pairs && copy_next < i + stages - 2; ++copy_next) VFM_TRY(enqueue_h2d(copy_next));
      if (i < n_pairs) VFM_TRY(stage1(i));
      if (i >= 1 && i - 1 < n_pairs) VFM_TRY(stage2(i - 1, nullptr));
      if (i >= 2) VFM_TRY(stage3(i - 2, nullptr));
    }
  } else {
    // GROUPED schedule (VFMREG_SCHEDULE=grouped, or fewer than 3 lanes): groups of `lanes` pairs, stage by stage; a search
    // kernel never shares the GPU with a small kernel (it then runs at the rate it has alone), the small kernels of a
    // group overlap each other.
    for (int g0 = 0; g0 < n_pairs; g0 += lanes) {
      const int gn = (n_pairs - g0) < lanes ? (n_pairs - g0) : lanes;
      if (host)   // this group's and the next group's uploads: pair j reuses the stage of pair j - stages, whose group is enqueued
        for (; copy_next < n_pairs && copy_next < g0 + stages; ++copy_next) VFM_TRY(enqueue_h2d(copy_next));
      if (ks && g0 > 0)   // the previous group's stage 3 has the GPU to itself
        for (int l = 0; l < lanes; ++l) VFM_CUDA(cudaStreamWaitEvent(ks, ctx->ev_join[l], 0));
      cudaEvent_t searches_done = nullptr;
      for (int k = 0; k < gn; ++k) VFM_TRY(stage1(g0 + k));
      if (ks) {
        searches_done = next_event();
        VFM_CUDA(cudaEventRecord(searches_done, ks));
      }
      for (int k = 0; k < gn; ++k) VFM_TRY(stage2(g0 + k, searches_done));
      if (ks) {
        searches_done = next_event();
        VFM_CUDA(cudaEventRecord(searches_done, ks));
      }
      for (int k = 0; k < gn; ++k) VFM_TRY(stage3(g0 + k, searches_done));
    }
  }
  ctx->stream = guard.saved;
  ctx->arena.limit = 0;
  ctx->match_stream = nullptr;   // every search was handed back to its lane by an event
  for (int l = 1; l < lanes; ++l) VFM_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[l], 0));
  VFM_CUDA(cudaMemcpyAsync(ctx->pinned, slots, slots_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  VFM_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n_pairs; ++i) fill_result(results + i, static_cast<const char*>(ctx->pinned) + (size_t)i * 256, params->inlier_thresh);
  return VFMREG_OK;
}

static int batch_device(vfmreg_ctx* ctx, const BatchView& b, const vfmreg_register_params* params, vfmreg_register_result* results) {
  return batch_run(ctx, b, params, results, false);
}

static int batch_host(vfmreg_ctx* ctx, const BatchView& b, const vfmreg_register_params* params, vfmreg_register_result* results) {
  return batch_run(ctx, b, params, results, true);
}

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
  hostcopy_destroy(ctx);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i < 2; ++i) {
      cudaEventDestroy(ctx->ev_ready[i]);
      cudaEventDestroy(ctx->ev_consumed[i]);
      cudaEventDestroy(ctx->ev_target_free[i]);
    }
  }
  for (int l = 1; l < vfmreg_ctx::MAX_LANES; ++l) {
    if (!ctx->lane_stream[l]) continue;
    cudaStreamSynchronize(ctx->lane_stream[l]);
    cudaStreamDestroy(ctx->lane_stream[l]);
    cudaEventDestroy(ctx->ev_join[l]);
  }
  if (ctx->ev_fork) {
    cudaEventDestroy(ctx->ev_fork);
    cudaEventDestroy(ctx->ev_join[0]);
  }
  if (ctx->match_stream_owned) {
    cudaStreamSynchronize(ctx->match_stream_owned);
    cudaStreamDestroy(ctx->match_stream_owned);
  }
  for (int i = 0; i < vfmreg_ctx::MATCH_EVENTS; ++i)
    if (ctx->match_ev[i]) cudaEventDestroy(ctx->match_ev[i]);
  for (int i = 0; i < ctx->map_ev_cap; ++i)
    if (ctx->map_ev[i]) cudaEventDestroy(ctx->map_ev[i]);
  free(ctx->map_ev);
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
  VFM_TRY(arena_reserve(ctx, prepared_bytes(n, d, flags) + prepared_bytes(m, d, flags) + search_scratch(ctx, n, m, d, flags, false) + 4096));
  Prepared A, B;
  VFM_TRY(prepare_arena(ctx, a, n, d, flags, &A));
  VFM_TRY(prepare_arena(ctx, b, m, d, flags, &B));
  return search_full(ctx, A, B, flags, idx01, sim01, sec01, idx10, sim10, sec10);
}

int vfmreg_filter_correspondences(vfmreg_ctx* ctx, const int32_t* idx01, const float* sim01, const float* sec01,
                                  const int32_t* idx10, int64_t n, float min_cos, float ratio, int mutual,
                                  int32_t* corr, int32_t* count) {
  VFM_CHECK_ARG(ctx && idx01 && sim01 && corr && count, "filter_correspondences: null pointer");
  VFM_CHECK_ARG(n >= 0, "filter_correspondences: negative n");
  VFM_CUDA(cudaSetDevice(ctx->device));
  return filter_corr(ctx, idx01, sim01, sec01, idx10, n, min_cos, ratio, mutual, corr, count);
}

int vfmreg_select_smallest(vfmreg_ctx* ctx, const int32_t* idx01, const float* sim01, int64_t n, int64_t n_keep, int32_t* corr,
                           float* dist, int32_t* count) {
  VFM_CHECK_ARG(ctx && idx01 && sim01 && corr && count, "select_smallest: null pointer");
  VFM_CHECK_ARG(n >= 0 && n_keep >= 0, "select_smallest: negative size");
  VFM_CUDA(cudaSetDevice(ctx->device));
  arena_reset(ctx);
  VFM_TRY(arena_reserve(ctx, arena_bytes((size_t)n + 1, 4) + 4096));
  return select_top(ctx, idx01, sim01, n, n_keep, corr, dist, count);
}

int vfmreg_l2_distances(vfmreg_ctx* ctx, const int32_t* idx01, const float* sim01, int64_t n, float* dist) {
  VFM_CHECK_ARG(ctx && sim01 && dist, "l2_distances: null pointer");
  VFM_CHECK_ARG(n >= 0, "l2_distances: negative n");
  VFM_CUDA(cudaSetDevice(ctx->device));
  return l2_from_sim(ctx, idx01, sim01, n, dist);
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

int vfmreg_register(vfmreg_ctx* ctx, const float* src_xyz, const float* tgt_xyz, const float* src_feats,
                    const float* tgt_feats, int64_t n, int64_t m, int32_t d, const vfmreg_register_params* params,
                    const int32_t* sample_idx, int32_t* corr_out, uint8_t* mask_out, vfmreg_register_result* result) {
  const BatchView b{1, &src_xyz, &src_feats, &n, &tgt_xyz, &tgt_feats, &m, nullptr, d, sample_idx ? &sample_idx : nullptr,
                    corr_out ? &corr_out : nullptr, mask_out ? &mask_out : nullptr};
  return batch_device(ctx, b, params, result);
}

int vfmreg_register_host(vfmreg_ctx* ctx, const float* src_xyz, const float* tgt_xyz, const float* src_feats,
                         const float* tgt_feats, int64_t n, int64_t m, int32_t d, const vfmreg_register_params* params,
                         const int32_t* sample_idx, int32_t* corr_out, uint8_t* mask_out,
                         vfmreg_register_result* result) {
  const BatchView b{1, &src_xyz, &src_feats, &n, &tgt_xyz, &tgt_feats, &m, nullptr, d, sample_idx ? &sample_idx : nullptr,
                    corr_out ? &corr_out : nullptr, mask_out ? &mask_out : nullptr};
  return batch_host(ctx, b, params, result);
}

int vfmreg_register_batch_host(vfmreg_ctx* ctx, int32_t n_pairs, const float* const* src_xyz, const float* const* tgt_xyz,
                               const float* const* src_feats, const float* const* tgt_feats, const int64_t* n, const int64_t* m,
                               int32_t d, const vfmreg_register_params* params, const int32_t* const* sample_idx,
                               int32_t* const* corr_out, uint8_t* const* mask_out, vfmreg_register_result* results) {
  const BatchView b{n_pairs, src_xyz, src_feats, n, tgt_xyz, tgt_feats, m, nullptr, d, sample_idx, corr_out, mask_out};
  return batch_host(ctx, b, params, results);
}

int vfmreg_register_batch(vfmreg_ctx* ctx, int32_t n_pairs, const float* const* src_xyz, const float* const* tgt_xyz,
                          const float* const* src_feats, const float* const* tgt_feats, const int64_t* n, const int64_t* m,
                          int32_t d, const vfmreg_register_params* params, const int32_t* const* sample_idx,
                          int32_t* const* corr_out, uint8_t* const* mask_out, vfmreg_register_result* results) {
  const BatchView b{n_pairs, src_xyz, src_feats, n, tgt_xyz, tgt_feats, m, nullptr, d, sample_idx, corr_out, mask_out};
  return batch_device(ctx, b, params, results);
}

// ---- resident maps ---------------------------------------------------------------------------------------------------
int vfmreg_map_create(vfmreg_ctx* ctx, const float* tgt_xyz, const float* tgt_feats, int64_t m, int32_t d, uint32_t flags,
                      int32_t host_buffers, vfmreg_map** out) {
  VFM_CHECK_ARG(ctx && out, "map_create: null pointer");
  *out = nullptr;
  VFM_CHECK_ARG(tgt_xyz && tgt_feats, "map_create: null pointer");
  VFM_CHECK_ARG(m > 0 && m < (1LL << 30) && d > 0, "map_create: empty map (m=%lld d=%d)", (long long)m, d);
  VFM_CUDA(cudaSetDevice(ctx->device));
  flags &= (VFMREG_NORMALIZE | VFMREG_ALGO_MASK);
  const size_t b_xyz = arena_bytes((size_t)m * 3, 4), b_prep = prepared_bytes(m, d, flags);
  vfmreg_map* map = new vfmreg_map();
  cudaError_t e = cudaMalloc(&map->slab, b_xyz + b_prep);
  if (e != cudaSuccess) {
    delete map;
    set_error("map_create: cudaMalloc(%zu) failed: %s", b_xyz + b_prep, cudaGetErrorString(e));
    return VFMREG_ERR_ALLOC;
  }
  map->ctx = ctx;
  map->m = m;
  map->d = d;
  map->flags = flags;
  map->xyz = reinterpret_cast<float*>(map->slab);
  int rc = VFMREG_OK;
  const float* feats_dev = tgt_feats;
  do {
    if (host_buffers) {
      // raw descriptors pass through the scratch arena on their way to the prepared rows
      arena_reset(ctx);
      if ((rc = arena_reserve(ctx, arena_bytes((size_t)m * d, 4))) != VFMREG_OK) break;
      float* raw = arena_take<float>(ctx, (size_t)m * d);
      if ((rc = h2d_copy(ctx, raw, tgt_feats, (size_t)m * d * 4, ctx->stream)) != VFMREG_OK ||
          (rc = h2d_copy(ctx, map->xyz, tgt_xyz, (size_t)m * 3 * 4, ctx->stream)) != VFMREG_OK)
        break;
      feats_dev = raw;
    } else if (cudaMemcpyAsync(map->xyz, tgt_xyz, (size_t)m * 3 * 4, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) {
      set_error("map_create: device copy failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = VFMREG_ERR_CUDA;
      break;
    }
    if ((rc = prepare_into(ctx, feats_dev, m, d, flags, map->slab + b_xyz, &map->prep)) != VFMREG_OK) break;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {   // the map may be used from any stream afterwards
      set_error("map_create: %s", cudaGetErrorString(cudaGetLastError()));
      rc = VFMREG_ERR_CUDA;
    }
  } while (0);
  if (rc != VFMREG_OK) {
    cudaFree(map->slab);
    delete map;
    return rc;
  }
  *out = map;
  return VFMREG_OK;
}

void vfmreg_map_destroy(vfmreg_map* map) {
  if (!map) return;
  cudaSetDevice(map->ctx->device);
  cudaDeviceSynchronize();
  cudaFree(map->slab);
  delete map;
}

int64_t vfmreg_map_size(const vfmreg_map* map) { return map ? map->m : 0; }

int vfmreg_map_match(vfmreg_ctx* ctx, const vfmreg_map* map, const float* queries, int64_t n, float min_cos, int32_t* idx01,
                     float* sim01, float* sec01) {
  VFM_CHECK_ARG(ctx && map && queries && idx01, "map_match: null pointer");
  VFM_CHECK_ARG(map->ctx == ctx, "map_match: the map belongs to another context");
  VFM_CHECK_ARG(n > 0, "map_match: empty query set");
  VFM_CUDA(cudaSetDevice(ctx->device));
  arena_reset(ctx);
  VFM_TRY(arena_reserve(ctx, prepared_bytes(n, map->d, map->flags) + search_scratch(ctx, n, map->m, map->d, map->flags, false) + 4096));
  Prepared A;
  VFM_TRY(prepare_arena(ctx, queries, n, map->d, map->flags, &A));
  return search_full(ctx, A, map->prep, map->flags, idx01, sim01, sec01, nullptr, nullptr, nullptr, sec01 ? NAN : min_cos);
}

int vfmreg_register_scans(vfmreg_ctx* ctx, const vfmreg_map* map, int32_t n_scans, const float* const* src_xyz,
                          const float* const* src_feats, const int64_t* n, const vfmreg_register_params* params,
                          const int32_t* const* sample_idx, int32_t host_buffers, int32_t* const* corr_out, uint8_t* const* mask_out,
                          vfmreg_register_result* results) {
  VFM_CHECK_ARG(ctx && map, "register_scans: null pointer");
  VFM_CHECK_ARG(params && ((params->flags ^ map->flags) & (VFMREG_NORMALIZE | VFMREG_ALGO_MASK)) == 0,
                "register_scans: the map was prepared with other NORMALIZE / ALGO flags");
  const BatchView b{n_scans, src_xyz, src_feats, n, nullptr, nullptr, nullptr, map, map->d, sample_idx, corr_out, mask_out};
  return host_buffers ? batch_host(ctx, b, params, results) : batch_device(ctx, b, params, results);
}

}  // extern "C"
