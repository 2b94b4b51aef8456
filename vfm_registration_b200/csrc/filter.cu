// Cosine gate + mutual check + ratio test -> ordered correspondence list (reference VoxelHashMap.cpp:501-511,
// 587-600 and registration_node.py:530).  One CTA, ballot-based ordered stream compaction: n is at most a few
// 10^4 queries, so this is a latency-sized kernel (reads 12-16 B, writes <= 8 B per query).
#include "common.cuh"

namespace vfm {

__global__ void __launch_bounds__(1024) filter_corr_kernel(const int32_t* __restrict__ idx01, const float* __restrict__ sim01,
                                                          const float* __restrict__ sec01, const int32_t* __restrict__ idx10,
                                                          int n, float min_cos, float ratio2, int use_cos, int use_ratio,
                                                          int mutual, int32_t* __restrict__ corr, int32_t* __restrict__ count) {
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t == 0) base_s = 0;
  __syncthreads();
  for (int start = 0; start < n; start += 1024) {
    const int i = start + t;
    bool keep = false;
    int j = -1;
    if (i < n) {
      j = idx01[i];
      keep = j >= 0;
      if (keep && use_cos) keep = sim01[i] >= min_cos;
      if (keep && mutual) keep = idx10[j] == i;
      if (keep && use_ratio) keep = __fsub_rn(1.0f, sim01[i]) < __fmul_rn(ratio2, __fsub_rn(1.0f, sec01[i]));
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_tot[w] = __popc(bal);
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll 1
    for (int k = 0; k < 32; ++k) {
      const int v = warp_tot[k];
      if (k < w) off += v;
      tot += v;
    }
    const int base = base_s;
    if (keep) {
      const int pos = base + off + __popc(bal & ((1u << lane) - 1u));
      corr[2 * pos] = i;
      corr[2 * pos + 1] = j;
    }
    __syncthreads();
    if (t == 0) base_s = base + tot;
    __syncthreads();
  }
  if (t == 0) *count = base_s;
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

}  // namespace vfm
