// Cosine gate + mutual check + ratio test -> ordered correspondence list (reference VoxelHashMap.cpp:501-511,
// 587-600 and registration_node.py:530).  One CTA: every thread owns a run of consecutive queries, counts its keepers, one
// block-wide scan gives its write position, a second pass writes them in order.  n is at most a few 10^4 queries, so this
// is a latency-sized kernel (reads 12-16 B, writes <= 8 B per query); the pruned mutual check (gate -> gather the map rows
// the gated queries point at -> reverse search over those rows -> keep the pairs that point back) lives here too.
#include "common.cuh"

namespace vfm {

// exclusive prefix of one int per thread over a 1024-thread CTA; *total = sum
__device__ __forceinline__ int block_exclusive_scan_1024(int v, int* wsum, int& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += u;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  if (w == 0) {
    int x = wsum[lane];
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

__device__ __forceinline__ bool corr_keep(const int32_t* __restrict__ idx01, const float* __restrict__ sim01, const float* __restrict__ sec01,
                                          const int32_t* __restrict__ idx10, int i, float min_cos, float ratio2, int use_cos,
                                          int use_ratio, int mutual, int& j) {
  j = idx01[i];
  bool keep = j >= 0;
  if (keep && use_cos) keep = sim01[i] >= min_cos;
  if (keep && mutual) keep = idx10[j] == i;
  if (keep && use_ratio) keep = __fsub_rn(1.0f, sim01[i]) < __fmul_rn(ratio2, __fsub_rn(1.0f, sec01[i]));
  return keep;
}

// Thread t owns the consecutive queries [t * per, (t + 1) * per): count, one block-wide scan, write -- two passes over
// at most a few 10^4 queries, three barriers in total.
__global__ void __launch_bounds__(1024) filter_corr_kernel(const int32_t* __restrict__ idx01, const float* __restrict__ sim01,
                                                          const float* __restrict__ sec01, const int32_t* __restrict__ idx10,
                                                          int n, float min_cos, float ratio2, int use_cos, int use_ratio,
                                                          int mutual, int32_t* __restrict__ corr, int32_t* __restrict__ count) {
  __shared__ int wsum[32];
  const int per = (n + 1023) / 1024;
  const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
  int mine = 0, j;
  for (int i = lo; i < hi; ++i) mine += corr_keep(idx01, sim01, sec01, idx10, i, min_cos, ratio2, use_cos, use_ratio, mutual, j) ? 1 : 0;
  int total;
  int pos = block_exclusive_scan_1024(mine, wsum, total);
  for (int i = lo; i < hi; ++i)
    if (corr_keep(idx01, sim01, sec01, idx10, i, min_cos, ratio2, use_cos, use_ratio, mutual, j)) {
      corr[2 * pos] = i;
      corr[2 * pos + 1] = j;
      ++pos;
    }
  if (threadIdx.x == 0) *count = total;
}

int filter_corr(vfmreg_ctx* ctx, const int32_t* idx01, const float* sim01, const float* sec01, const int32_t* idx10,
                int64_t n, float min_cos, float ratio, int mutual, int32_t* corr, int32_t* count) {
  const int use_cos = !(min_cos != min_cos);
  const int use_ratio = !(ratio != ratio);
  VFM_CHECK_ARG(!use_ratio || sec01, "ratio test needs sec01");
  VFM_CHECK_ARG(!mutual || idx10, "mutual filter needs idx10");
  VFM_CHECK_ARG(n < (1LL << 31), "n too large");
  filter_corr_kernel<<<1, 1024, 0, ctx->stream>>>(idx01, sim01, sec01, idx10, (int)n, min_cos, ratio * ratio, use_cos,
                                                 use_ratio, mutual, corr, count);
  return launch_check(ctx, "filter_corr");
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

__global__ void __launch_bounds__(1024) filter_mutual_list_kernel(const int32_t* __restrict__ cand, const int32_t* __restrict__ cand_count,
                                                                 const int32_t* __restrict__ back, int max_rows,
                                                                 int32_t* __restrict__ corr, int32_t* __restrict__ count) {
  __shared__ int wsum[32];
  const int n = min(*cand_count, max_rows);
  const int per = (n + 1023) / 1024;
  const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
  int mine = 0;
  for (int k = lo; k < hi; ++k) mine += (back[k] == cand[2 * k]) ? 1 : 0;
  int total;
  int pos = block_exclusive_scan_1024(mine, wsum, total);
  for (int k = lo; k < hi; ++k) {
    const int i = cand[2 * k];
    if (back[k] == i) {
      corr[2 * pos] = i;
      corr[2 * pos + 1] = cand[2 * k + 1];
      ++pos;
    }
  }
  if (threadIdx.x == 0) *count = total;
}

int filter_mutual_list(vfmreg_ctx* ctx, const int32_t* cand, const int32_t* cand_count, const int32_t* back, int64_t max_rows,
                       int32_t* corr, int32_t* count) {
  filter_mutual_list_kernel<<<1, 1024, 0, ctx->stream>>>(cand, cand_count, back, (int)max_rows, corr, count);
  return launch_check(ctx, "filter_mutual_list");
}

}  // namespace vfm
