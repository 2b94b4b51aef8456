// Cosine gate + mutual check + ratio test -> ordered correspondence list (reference VoxelHashMap.cpp:501-511,
// 587-600 and registration_node.py:530).  One CTA per list (ordered_compact below).  n is at most a few 10^4 queries, so
// this is a latency-sized kernel (reads 12-16 B, writes <= 8 B per query); the pruned mutual check (gate -> gather the map
// rows the gated queries point at -> reverse search over those rows -> keep the pairs that point back) lives here too.
#include "common.cuh"

namespace vfm {

// One-CTA ordered compaction.  FILTER_THREADS = 512 threads (16 warps x 32 registers: small enough to be placed at once
// beside the candidate-search CTA of a neighbouring pair -- these kernels sit on the latency chain between a pair's two
// searches).  The index range is walked in chunks of CHUNK_TILES tiles of 512 consecutive indices: every thread first
// evaluates its element of each tile (coalesced, independent loads -> one memory round trip per chunk) into a bit mask,
// one block-wide scan over the (tile, warp) counts gives every warp its write position, then the kept elements are written
// in index order.
constexpr int FILTER_THREADS = 512;
constexpr int FILTER_WARPS = FILTER_THREADS / 32;
constexpr int CHUNK_TILES = 32;
static_assert(CHUNK_TILES * FILTER_WARPS == FILTER_THREADS, "one scan entry per thread");

// exclusive prefix of one int per thread over the CTA; total = sum (all threads)
__device__ __forceinline__ int block_exclusive_scan(int v, int* wsum, int& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += u;
  }
  __syncthreads();   // wsum may still be read by the previous call
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  if (w == 0) {
    int x = lane < nw ? wsum[lane] : 0;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += u;
    }
    wsum[lane] = x;
  }
  __syncthreads();
  total = wsum[31];
  return (w > 0 ? wsum[w - 1] : 0) + incl - v;
}

// keep(i) -> bool, emit(i, pos) writes element i at output position pos; returns the number of kept elements
template <class Keep, class Emit>
__device__ __forceinline__ int ordered_compact(int n, Keep keep, Emit emit) {
  __shared__ int wsum[32];
  __shared__ int wbase[CHUNK_TILES * FILTER_WARPS];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  int running = 0;
  for (int c0 = 0; c0 < n; c0 += CHUNK_TILES * FILTER_THREADS) {
    uint32_t mask = 0;
#pragma unroll 8
    for (int r = 0; r < CHUNK_TILES; ++r) {
      const int i = c0 + r * FILTER_THREADS + t;
      if (i < n && keep(i)) mask |= 1u << r;
    }
    __syncthreads();   // wbase of the previous chunk has been consumed
#pragma unroll
    for (int r = 0; r < CHUNK_TILES; ++r) {
      const unsigned b = __ballot_sync(0xffffffffu, (mask >> r) & 1u);
      if (lane == 0) wbase[r * FILTER_WARPS + w] = __popc(b);
    }
    __syncthreads();
    int total;
    const int excl = block_exclusive_scan(wbase[t], wsum, total);   // entry t = (tile t / 16, warp t % 16): index order
    __syncthreads();
    wbase[t] = running + excl;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < CHUNK_TILES; ++r) {
      const bool mine = (mask >> r) & 1u;
      const unsigned b = __ballot_sync(0xffffffffu, mine);
      if (mine) emit(c0 + r * FILTER_THREADS + t, wbase[r * FILTER_WARPS + w] + __popc(b & ((1u << lane) - 1u)));
    }
    running += total;
  }
  return running;
}

__device__ __forceinline__ bool corr_keep(const int32_t* __restrict__ idx01, const float* __restrict__ sim01, const float* __restrict__ sec01,
                                          const int32_t* __restrict__ idx10, int i, float min_cos, float ratio2, int use_cos,
                                          int use_ratio, int mutual) {
  const int j = idx01[i];
  bool keep = j >= 0;
  if (keep && use_cos) keep = sim01[i] >= min_cos;
  if (keep && mutual) keep = idx10[j] == i;
  if (keep && use_ratio) keep = __fsub_rn(1.0f, sim01[i]) < __fmul_rn(ratio2, __fsub_rn(1.0f, sec01[i]));
  return keep;
}

__global__ void __launch_bounds__(FILTER_THREADS) filter_corr_kernel(const int32_t* __restrict__ idx01, const float* __restrict__ sim01,
                                                                    const float* __restrict__ sec01, const int32_t* __restrict__ idx10,
                                                                    int n, float min_cos, float ratio2, int use_cos, int use_ratio,
                                                                    int mutual, int32_t* __restrict__ corr, int32_t* __restrict__ count) {
  const int total = ordered_compact(
      n, [&](int i) { return corr_keep(idx01, sim01, sec01, idx10, i, min_cos, ratio2, use_cos, use_ratio, mutual); },
      [&](int i, int pos) {
        corr[2 * pos] = i;
        corr[2 * pos + 1] = idx01[i];
      });
  if (threadIdx.x == 0) *count = total;
}

int filter_corr(vfmreg_ctx* ctx, const int32_t* idx01, const float* sim01, const float* sec01, const int32_t* idx10,
                int64_t n, float min_cos, float ratio, int mutual, int32_t* corr, int32_t* count) {
  const int use_cos = !(min_cos != min_cos);
  const int use_ratio = !(ratio != ratio);
  VFM_CHECK_ARG(!use_ratio || sec01, "ratio test needs sec01");
  VFM_CHECK_ARG(!mutual || idx10, "mutual filter needs idx10");
  VFM_CHECK_ARG(n < (1LL << 31), "n too large");
  filter_corr_kernel<<<1, FILTER_THREADS, 0, ctx->stream>>>(idx01, sim01, sec01, idx10, (int)n, min_cos, ratio * ratio, use_cos,
                                                           use_ratio, mutual, corr, count);
  return launch_check(ctx, "filter_corr");
}

// ---- "keep the n_points smallest distances" (registration_node.py:212-214 and :510-518) -------------------------------------
// The reference takes np.argpartition(dists, n)[:n] of the nearest-neighbour distances (an unordered set; ties at the
// boundary are whatever introselect leaves).  For unit vectors the distance sqrt(2 - 2 s + 1e-6) falls with the similarity s,
// so the set is the n largest similarities: a 4 x 8-bit radix select over the order-preserving integer image of s finds
// the n-th largest key, ties at the boundary are broken towards the lowest query index (deterministic), and the kept
// (query, match) pairs are emitted in query order.  Queries without a match (index < 0) never qualify.  One CTA.
__device__ __forceinline__ uint32_t order_key(float s) {
  const uint32_t u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(FILTER_THREADS) select_top_kernel(const int32_t* __restrict__ idx01, const float* __restrict__ sim01, int n,
                                                                   int n_keep, int32_t* __restrict__ eq_list,
                                                                   int32_t* __restrict__ corr, float* __restrict__ dist,
                                                                   int32_t* __restrict__ count) {
  __shared__ unsigned hist[256];
  __shared__ uint32_t s_prefix;
  __shared__ int s_need, s_valid;
  const int t = threadIdx.x;
  // how many queries have a match at all
  if (t == 0) s_valid = 0;
  __syncthreads();
  int mine = 0;
  for (int i = t; i < n; i += FILTER_THREADS) mine += idx01[i] >= 0 ? 1 : 0;
  if (mine) atomicAdd(&s_valid, mine);
  __syncthreads();
  const int valid = s_valid;
  const int want = min(max(n_keep, 0), valid);
  uint32_t thr_key = 0;   // keep key > thr_key, plus the first `need_eq` (by index) with key == thr_key
  int need_eq = 0;
  if (want > 0 && want < valid) {
    if (t == 0) {
      s_prefix = 0;
      s_need = want;
    }
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int bin = t; bin < 256; bin += FILTER_THREADS) hist[bin] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      for (int i = t; i < n; i += FILTER_THREADS) {
        if (idx01[i] < 0) continue;
        const uint32_t key = order_key(sim01[i]);
        if (pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&hist[(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (t == 0) {   // walk the bins from the largest digit down to the one that holds the s_need-th largest key
        int need = s_need, bin = 255;
        for (; bin > 0; --bin) {
          if ((int)hist[bin] >= need) break;
          need -= (int)hist[bin];
        }
        s_prefix = prefix | ((uint32_t)bin << shift);
        s_need = need;
      }
      __syncthreads();
    }
    thr_key = s_prefix;
    need_eq = s_need;
  }
  const bool all = (want == valid);
  // boundary ties: the indices of the queries whose key equals the threshold, in index order
  int cut = -1;   // equal keys are kept up to this query index
  if (!all && want > 0) {
    ordered_compact(
        n, [&](int i) { return idx01[i] >= 0 && order_key(sim01[i]) == thr_key; }, [&](int i, int pos) { eq_list[pos] = i; });
    __syncthreads();
    cut = eq_list[need_eq - 1];
  }
  const int total = ordered_compact(
      n,
      [&](int i) {
        if (idx01[i] < 0 || want == 0) return false;
        if (all) return true;
        const uint32_t key = order_key(sim01[i]);
        return key > thr_key || (key == thr_key && i <= cut);
      },
      [&](int i, int pos) {
        corr[2 * pos] = i;
        corr[2 * pos + 1] = idx01[i];
        if (dist) dist[pos] = __fsqrt_rn(__fadd_rn(__fsub_rn(2.0f, __fmul_rn(2.0f, sim01[i])), 1e-6f));
      });
  if (t == 0) *count = total;
}

int select_top(vfmreg_ctx* ctx, const int32_t* idx01, const float* sim01, int64_t n, int64_t n_keep, int32_t* corr, float* dist,
               int32_t* count) {
  VFM_CHECK_ARG(n >= 0 && n < (1LL << 31), "select_smallest: bad n");
  int32_t* eq_list = arena_take<int32_t>(ctx, (size_t)(n > 0 ? n : 1));
  if (!eq_list) {
    set_error("select_smallest: scratch arena too small");
    return VFMREG_ERR_ALLOC;
  }
  select_top_kernel<<<1, FILTER_THREADS, 0, ctx->stream>>>(idx01, sim01, (int)n, (int)(n_keep < n ? n_keep : n), eq_list, corr, dist, count);
  return launch_check(ctx, "select_top_kernel");
}

// dist[i] = sqrt(2 - 2 s_i + 1e-6): the L2 distance of unit vectors as the reference's brute-force block writes it
// (registration_node.py:197-198); queries without a match get +inf
__global__ void l2_from_sim_kernel(const int32_t* __restrict__ idx01, const float* __restrict__ sim01, int n, float* __restrict__ dist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dist[i] = (idx01 && idx01[i] < 0) ? INFINITY : __fsqrt_rn(__fadd_rn(__fsub_rn(2.0f, __fmul_rn(2.0f, sim01[i])), 1e-6f));
}

int l2_from_sim(vfmreg_ctx* ctx, const int32_t* idx01, const float* sim01, int64_t n, float* dist) {
  if (n <= 0) return VFMREG_OK;
  l2_from_sim_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(idx01, sim01, (int)n, dist);
  return launch_check(ctx, "l2_from_sim_kernel");
}

// ---- pruned mutual check -------------------------------------------------------------------------------------------
// idx10[j] is only ever read at j = idx01[i] of the queries that passed the gate, so the reverse search runs over those
// map rows only (a compacted copy) instead of all m; the answer per row is unchanged because rows are independent.

// one warp per listed row: 128-bit copies of the fp32 row, the fp16 row and the non-zero flag
__global__ void __launch_bounds__(256) gather_rows_kernel(const int32_t* __restrict__ pairs, const int32_t* __restrict__ count,
                                                          int max_rows, int col, int dp, const float* __restrict__ b32,
                                                          const uint16_t* __restrict__ b16, const uint8_t* __restrict__ nzb,
                                                          float* __restrict__ o32, uint16_t* __restrict__ o16,
                                                          uint8_t* __restrict__ onz, const float* __restrict__ sim,
                                                          float* __restrict__ osim) {
  const int rows = min(*count, max_rows);
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < rows; k += warps) {
    const int j = pairs[2 * k + col];
    const uint4* s32 = reinterpret_cast<const uint4*>(b32 + (size_t)j * dp);
    uint4* d32 = reinterpret_cast<uint4*>(o32 + (size_t)k * dp);
    for (int q = lane; q < dp / 4; q += 32) d32[q] = __ldg(s32 + q);
    const uint4* s16 = reinterpret_cast<const uint4*>(b16 + (size_t)j * dp);
    uint4* d16 = reinterpret_cast<uint4*>(o16 + (size_t)k * dp);
    for (int q = lane; q < dp / 8; q += 32) d16[q] = __ldg(s16 + q);
    if (lane == 0) {
      onz[k] = nzb[j];
      if (osim) osim[k] = sim[pairs[2 * k]];
    }
  }
}

int gather_rows(vfmreg_ctx* ctx, const int32_t* pairs, const int32_t* count, int64_t max_rows, int col, int dp, const float* b32,
                const void* b16, const uint8_t* nzb, float* o32, void* o16, uint8_t* onz, const float* sim, float* osim) {
  VFM_CHECK_ARG(dp % 8 == 0, "gather_rows: padded dim %d not a multiple of 8", dp);
  const int blocks = (int)((max_rows + 7) / 8 < ctx->sm_count * 8 ? (max_rows + 7) / 8 : ctx->sm_count * 8);
  gather_rows_kernel<<<blocks > 0 ? blocks : 1, 256, 0, ctx->stream>>>(pairs, count, (int)max_rows, col, dp, b32,
                                                                      static_cast<const uint16_t*>(b16), nzb, o32,
                                                                      static_cast<uint16_t*>(o16), onz, sim, osim);
  return launch_check(ctx, "gather_rows");
}

__global__ void __launch_bounds__(FILTER_THREADS) filter_mutual_list_kernel(const int32_t* __restrict__ cand, const int32_t* __restrict__ cand_count,
                                                                           const int32_t* __restrict__ back, int max_rows,
                                                                           int32_t* __restrict__ corr, int32_t* __restrict__ count) {
  const int n = min(*cand_count, max_rows);
  const int total = ordered_compact(
      n, [&](int k) { return back[k] == cand[2 * k]; },
      [&](int k, int pos) {
        corr[2 * pos] = cand[2 * k];
        corr[2 * pos + 1] = cand[2 * k + 1];
      });
  if (threadIdx.x == 0) *count = total;
}

int filter_mutual_list(vfmreg_ctx* ctx, const int32_t* cand, const int32_t* cand_count, const int32_t* back, int64_t max_rows,
                       int32_t* corr, int32_t* count) {
  filter_mutual_list_kernel<<<1, FILTER_THREADS, 0, ctx->stream>>>(cand, cand_count, back, (int)max_rows, corr, count);
  return launch_check(ctx, "filter_mutual_list");
}

}  // namespace vfm
