// Batched correspondence RANSAC (reference call site registration_node.py:319-327:
// o3d registration_ransac_based_on_correspondence, ransac_n = 3, PointToPoint(False), confidence 1 => every
// iteration runs).  float64 throughout, in the canonical operation order shared with oracle/c/oracle_ref.c so that
// hypotheses, inlier counts, masks and the returned transform are bit-identical to the CPU restatement.
//
// THIS FILE IS COMPILED WITH -fmad=false: every fused multiply-add is an explicit fma().
//
// Kernels
//   gather_kabsch_kernel  one launch: corr (K x 2) + xyz -> pq (K x 6) float64 (p.xyz, q.xyz), 48 B/correspondence; and one
//                      thread per hypothesis: 3 samples (given or counter-based RNG) -> R,t (12 doubles).
//   score_kernel       grid (hyp blocks) x (K slices); 256 hypotheses per CTA, two per thread, R,t in registers;
//                      correspondences staged through a 3 KB shared-memory tile (the kernel runs beside the candidate
//                      search of the neighbouring pairs, which leaves ~18 KB per SM); inlier count and the quantised residual sum
//                      (rint(d^2 2^40/tau^2), order-independent integers) are added with integer atomics.
//                      Bound: FP64 FMA pipe (16 flop-pairs per hypothesis x correspondence); HBM traffic is the
//                      48 K bytes of pq per hypothesis block, L2 resident.
//   finalize_kernel    one CTA: arg-best by (count desc, sumq asc, id asc) with warp-shuffle reductions, the winner's
//                      inlier mask, optional least-squares refit, outputs.
#include <stdlib.h>

#include "common.cuh"

namespace vfm {

constexpr int SCORE_THREADS = 128;
constexpr int PREP_THREADS = 128;
constexpr int SCORE_TILE = 64;   // correspondences per shared-memory tile (3 KB)
#ifndef VFM_HYP_PER_THREAD
#define VFM_HYP_PER_THREAD 2
#endif
#ifndef VFM_SCORE_MIN_CTAS
#define VFM_SCORE_MIN_CTAS 5
#endif
constexpr int HYP_PER_THREAD = VFM_HYP_PER_THREAD;   // hypotheses per thread: every shared-memory broadcast feeds that many DFMA chains
constexpr int HYP_PER_CTA = HYP_PER_THREAD * SCORE_THREADS;
constexpr int SWEEPS = 6;
constexpr int FIN_THREADS = 256;   // 256 x <= 128 registers: the CTA must fit beside a candidate-search CTA of a neighbouring pair

__device__ __forceinline__ uint32_t sample_u32(uint64_t seed, uint64_t ctr, uint32_t k) {
  uint64_t z = seed * 0x9E3779B97F4A7C15ULL + ctr;
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  return (uint32_t)(((z >> 32) * (uint64_t)k) >> 32);
}

// Rigid fit from the 3x3 cross-covariance S (row-major): see oracle_ref.c:orc_fit_from_sigma for the canonical
// order.  M = S^T S, cyclic Jacobi (6 sweeps), u_i = S v_i / sigma_i, one Gram-Schmidt step, third pair by cross
// products (= U diag(1,1,det U det V) V^T).  Returns false for rank-deficient samples.
__host__ __device__ bool fit_from_sigma(const double* S, const double* pm, const double* qm, double* rt) {
  double a[3][3], v[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      a[i][j] = (S[0 * 3 + i] * S[0 * 3 + j] + S[1 * 3 + i] * S[1 * 3 + j]) + S[2 * 3 + i] * S[2 * 3 + j];
      v[i][j] = (i == j) ? 1.0 : 0.0;
    }
#pragma unroll 1
  for (int sweep = 0; sweep < SWEEPS; ++sweep) {
#pragma unroll
    for (int e = 0; e < 3; ++e) {
      const int p = (e == 2) ? 1 : 0, q = (e == 0) ? 1 : 2, r = (e == 0) ? 2 : ((e == 1) ? 1 : 0);
      const double apq = a[p][q];
      if (apq == 0.0) continue;
      const double app = a[p][p], aqq = a[q][q];
      const double theta = (aqq - app) / (2.0 * apq);
      double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
      if (theta < 0.0) t = -t;
      const double c = 1.0 / sqrt(t * t + 1.0);
      const double s = t * c;
      a[p][p] = app - t * apq;
      a[q][q] = aqq + t * apq;
      a[p][q] = 0.0;
      a[q][p] = 0.0;
      const double arp = a[r][p], arq = a[r][q];
      const double nrp = c * arp - s * arq;
      const double nrq = s * arp + c * arq;
      a[r][p] = nrp;
      a[p][r] = nrp;
      a[r][q] = nrq;
      a[q][r] = nrq;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double vip = v[i][p], viq = v[i][q];
        v[i][p] = c * vip - s * viq;
        v[i][q] = s * vip + c * viq;
      }
    }
  }
  int i0 = 0, i1 = 1, i2 = 2, tmp;
  double l0 = a[0][0], l1 = a[1][1], l2 = a[2][2], lt;
  if (l1 > l0) { lt = l0; l0 = l1; l1 = lt; tmp = i0; i0 = i1; i1 = tmp; }
  if (l2 > l0) { lt = l0; l0 = l2; l2 = lt; tmp = i0; i0 = i2; i2 = tmp; }
  if (l2 > l1) { lt = l1; l1 = l2; l2 = lt; tmp = i1; i1 = i2; i2 = tmp; }
  if (!(l0 > 1e-300) || !(l1 > 1e-12 * l0)) {
#pragma unroll
    for (int i = 0; i < 12; ++i) rt[i] = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
    return false;
  }
  const double s1 = sqrt(l0), s2 = sqrt(l1);
  double v1[3], v2[3], v3[3], u1[3], u2[3], u3[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    v1[i] = (i0 == 0) ? v[i][0] : ((i0 == 1) ? v[i][1] : v[i][2]);
    v2[i] = (i1 == 0) ? v[i][0] : ((i1 == 1) ? v[i][1] : v[i][2]);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    u1[i] = ((S[i * 3 + 0] * v1[0] + S[i * 3 + 1] * v1[1]) + S[i * 3 + 2] * v1[2]) / s1;
    u2[i] = ((S[i * 3 + 0] * v2[0] + S[i * 3 + 1] * v2[1]) + S[i * 3 + 2] * v2[2]) / s2;
  }
  const double n1 = sqrt((u1[0] * u1[0] + u1[1] * u1[1]) + u1[2] * u1[2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) u1[i] = u1[i] / n1;
  const double dp = (u1[0] * u2[0] + u1[1] * u2[1]) + u1[2] * u2[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) u2[i] = u2[i] - dp * u1[i];
  const double n2 = sqrt((u2[0] * u2[0] + u2[1] * u2[1]) + u2[2] * u2[2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) u2[i] = u2[i] / n2;
  u3[0] = u1[1] * u2[2] - u1[2] * u2[1];
  u3[1] = u1[2] * u2[0] - u1[0] * u2[2];
  u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
  v3[0] = v1[1] * v2[2] - v1[2] * v2[1];
  v3[1] = v1[2] * v2[0] - v1[0] * v2[2];
  v3[2] = v1[0] * v2[1] - v1[1] * v2[0];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) rt[i * 3 + j] = (u1[i] * v1[j] + u2[i] * v2[j]) + u3[i] * v3[j];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    rt[9 + i] = qm[i] - ((rt[i * 3 + 0] * pm[0] + rt[i * 3 + 1] * pm[1]) + rt[i * 3 + 2] * pm[2]);
  return true;
}

// the same routine for host callers (teaser.cu: weighted Kabsch of the GNC-TLS rotation loop)
bool host_fit_from_sigma(const double* S, const double* pm, const double* qm, double* rt) { return fit_from_sigma(S, pm, qm, rt); }

__device__ __forceinline__ double resid2(const double* rt, double px, double py, double pz, double qx, double qy,
                                         double qz) {
  const double ex = fma(rt[0], px, fma(rt[1], py, fma(rt[2], pz, rt[9])));
  const double ey = fma(rt[3], px, fma(rt[4], py, fma(rt[5], pz, rt[10])));
  const double ez = fma(rt[6], px, fma(rt[7], py, fma(rt[8], pz, rt[11])));
  const double dx = ex - qx, dy = ey - qy, dz = ez - qz;
  return fma(dx, dx, fma(dy, dy, dz * dz));
}

__device__ __forceinline__ long long quant(double d2, double scale) {
  const double z = d2 * scale + 4503599627370496.0;  // 2^52: integer lands in the mantissa, round-to-nearest-even
  return __double_as_longlong(z) - 0x4330000000000000LL;
}

// One launch for the two steps that only depend on the correspondence list: blocks [0, gather_blocks) copy the listed points
// into pq (K x 6 float64, what score_kernel streams); the other blocks fit one hypothesis per thread from its 3 samples,
// reading the sampled points through corr directly (the same float64 values pq receives, so neither half waits for the other).
template <typename XYZ>
__global__ void __launch_bounds__(PREP_THREADS)
    gather_kabsch_kernel(const XYZ* __restrict__ src, const XYZ* __restrict__ tgt, const int32_t* __restrict__ corr,
                         const int32_t* __restrict__ count, int max_corr, int gather_blocks, double* __restrict__ pq,
                         const int32_t* __restrict__ sample_idx, int n_hyp, uint64_t seed, double* __restrict__ rts,
                         int32_t* __restrict__ counts, unsigned long long* __restrict__ sumq) {
  const int K = min(*count, max_corr);
  if ((int)blockIdx.x < gather_blocks) {
    const int k = blockIdx.x * PREP_THREADS + threadIdx.x;
    if (k >= K) return;
    const int i = corr[2 * k], j = corr[2 * k + 1];
    double* o = pq + (int64_t)k * 6;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o[c] = (double)src[(int64_t)i * 3 + c];
      o[3 + c] = (double)tgt[(int64_t)j * 3 + c];
    }
    return;
  }
  const int h = (blockIdx.x - gather_blocks) * PREP_THREADS + threadIdx.x;
  if (h >= n_hyp) return;
  if (sumq) sumq[h] = 0ULL;
  double* rt = rts + (int64_t)h * 12;
  if (K < 3) {
    counts[h] = -1;
    return;
  }
  double p[9], q[9];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    int s = sample_idx ? sample_idx[h * 3 + j] : (int)sample_u32(seed, (uint64_t)h * 3 + j, (uint32_t)K);
    s = min(max(s, 0), K - 1);
    const int ci = corr[2 * s], cj = corr[2 * s + 1];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      p[j * 3 + c] = (double)src[(int64_t)ci * 3 + c];
      q[j * 3 + c] = (double)tgt[(int64_t)cj * 3 + c];
    }
  }
  double pm[3], qm[3], S[9], out[12];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    pm[i] = ((p[0 + i] + p[3 + i]) + p[6 + i]) / 3.0;
    qm[i] = ((q[0 + i] + q[3 + i]) + q[6 + i]) / 3.0;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      S[i * 3 + j] = ((q[0 + i] - qm[i]) * (p[0 + j] - pm[j]) + (q[3 + i] - qm[i]) * (p[3 + j] - pm[j])) +
                     (q[6 + i] - qm[i]) * (p[6 + j] - pm[j]);
  const bool ok = fit_from_sigma(S, pm, qm, out);
#pragma unroll
  for (int i = 0; i < 12; ++i) rt[i] = out[i];
  counts[h] = ok ? 0 : -1;
}

// Two hypotheses per thread, R,t of both in registers; the CTA's slice of the correspondence list goes through a 3 KB
// shared-memory tile (64 rows, coalesced 128-bit loads) and is read back as warp-wide broadcasts serving both hypotheses.
// The tile is that small on purpose: the kernel runs beside the candidate search of the neighbouring pairs, whose CTAs
// leave about 18 KB of shared memory (and hardly any L1) per SM -- four of these CTAs still fit there.
// grid = (blocks of 256 hypotheses) x (slices of the correspondence list).
__global__ void __launch_bounds__(SCORE_THREADS, VFM_SCORE_MIN_CTAS)
    score_kernel(const double* __restrict__ pq, const int32_t* __restrict__ count, int max_corr,
                 const double* __restrict__ rts, int n_hyp, double tau2, double scale, int32_t* __restrict__ counts,
                 unsigned long long* __restrict__ sumq) {
  __shared__ __align__(16) double tile[SCORE_TILE * 6];
  const int K = min(*count, max_corr);
  if (K < 3) return;
  const int per = (K + gridDim.y - 1) / gridDim.y;
  const int k0 = blockIdx.y * per, k1 = min(K, k0 + per);
  if (k0 >= k1) return;
  // hypothesis j of this thread: blockIdx.x * HYP_PER_CTA + j * SCORE_THREADS + threadIdx.x
  int hid[HYP_PER_THREAD];
  bool live[HYP_PER_THREAD];
  double r[HYP_PER_THREAD][12];
  int cnt[HYP_PER_THREAD];
  long long sum[HYP_PER_THREAD];
#pragma unroll
  for (int j = 0; j < HYP_PER_THREAD; ++j) {
    hid[j] = blockIdx.x * HYP_PER_CTA + j * SCORE_THREADS + threadIdx.x;
    // counts[h] is only ever raised by score CTAs, never below 0: -1 marks a degenerate sample
    live[j] = (hid[j] < n_hyp) && (counts[hid[j]] >= 0);
#pragma unroll
    for (int i = 0; i < 12; ++i) r[j][i] = (hid[j] < n_hyp) ? rts[(int64_t)hid[j] * 12 + i] : 0.0;
    cnt[j] = 0;
    sum[j] = 0;
  }
  const double2* s2 = reinterpret_cast<const double2*>(tile);
  for (int t0 = k0; t0 < k1; t0 += SCORE_TILE) {
    const int rows = min(SCORE_TILE, k1 - t0);
    __syncthreads();
    const double2* g = reinterpret_cast<const double2*>(pq + (int64_t)t0 * 6);
    for (int i = threadIdx.x; i < rows * 3; i += SCORE_THREADS) reinterpret_cast<double2*>(tile)[i] = __ldg(g + i);
    __syncthreads();
#pragma unroll 2
    for (int k = 0; k < rows; ++k) {
      const double2 v0 = s2[3 * k], v1 = s2[3 * k + 1], v2 = s2[3 * k + 2];  // warp-wide broadcasts, shared by the thread's hypotheses
#pragma unroll
      for (int j = 0; j < HYP_PER_THREAD; ++j) {
        const double d = resid2(r[j], v0.x, v0.y, v1.x, v1.y, v2.x, v2.y);
        if (d < tau2) {
          ++cnt[j];
          sum[j] += quant(d, scale);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < HYP_PER_THREAD; ++j)
    if (live[j] && cnt[j] > 0) {
      atomicAdd(&counts[hid[j]], cnt[j]);
      atomicAdd(&sumq[hid[j]], (unsigned long long)sum[j]);
    }
}

struct Best {
  int count;
  unsigned long long sumq;
  int id;
};

__device__ __forceinline__ bool better(const Best& y, const Best& x) {  // is y better than x
  if (y.count != x.count) return y.count > x.count;
  if (y.count < 0) return false;
  if (y.sumq != x.sumq) return y.sumq < x.sumq;
  return y.id < x.id;
}

__global__ void __launch_bounds__(FIN_THREADS, 2)
    finalize_kernel(const double* __restrict__ pq, const int32_t* __restrict__ count, int max_corr,
                    const double* __restrict__ rts, int n_hyp, double tau2, int refit, const int32_t* __restrict__ counts,
                    const unsigned long long* __restrict__ sumq, double* __restrict__ T, int32_t* __restrict__ counts_out,
                    int64_t* __restrict__ sumq_out, uint8_t* __restrict__ mask, int64_t* __restrict__ stats) {
  __shared__ Best wbest[32];
  __shared__ double red[FIN_THREADS / 32][16];
  __shared__ double rt_s[12];
  __shared__ int best_s;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int K = min(*count, max_corr);
  Best b = {-1, 0ULL, 0x7fffffff};
  for (int h = t; h < n_hyp; h += FIN_THREADS) {
    const Best c = {counts[h], sumq[h], h};
    if (counts_out) counts_out[h] = c.count;
    if (sumq_out) sumq_out[h] = (int64_t)c.sumq;
    if (c.count >= 0 && better(c, b)) b = c;
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    Best o;
    o.count = __shfl_xor_sync(0xffffffffu, b.count, off);
    o.sumq = __shfl_xor_sync(0xffffffffu, b.sumq, off);
    o.id = __shfl_xor_sync(0xffffffffu, b.id, off);
    if (better(o, b)) b = o;
  }
  if (lane == 0) wbest[w] = b;
  __syncthreads();
  if (w == 0) {
    b = (lane < FIN_THREADS / 32) ? wbest[lane] : Best{-1, 0ULL, 0x7fffffff};
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      Best o;
      o.count = __shfl_xor_sync(0xffffffffu, b.count, off);
      o.sumq = __shfl_xor_sync(0xffffffffu, b.sumq, off);
      o.id = __shfl_xor_sync(0xffffffffu, b.id, off);
      if (better(o, b)) b = o;
    }
    if (lane == 0) {
      best_s = (b.count >= 0) ? b.id : -1;
      if (b.count >= 0)
        for (int i = 0; i < 12; ++i) rt_s[i] = rts[(int64_t)b.id * 12 + i];
      stats[0] = (b.count >= 0) ? b.id : -1;
      stats[1] = (b.count >= 0) ? b.count : 0;
      stats[2] = (b.count >= 0) ? (int64_t)b.sumq : 0;
      stats[3] = K;
    }
  }
  __syncthreads();
  const int best = best_s;
  if (best < 0) {
    if (t < 16) T[t] = (t % 5 == 0) ? 1.0 : 0.0;
    if (mask)
      for (int k = t; k < max_corr; k += FIN_THREADS) mask[k] = 0;
    return;
  }
  double rt[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) rt[i] = rt_s[i];
  // winner's inlier mask (+ sums for the optional refit: n, sum p, sum q, sum q p^T)
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.0;
  for (int k = t; k < max_corr; k += FIN_THREADS) {
    bool in = false;
    if (k < K) {
      const double* r = pq + (int64_t)k * 6;
      const double px = r[0], py = r[1], pz = r[2], qx = r[3], qy = r[4], qz = r[5];
      in = resid2(rt, px, py, pz, qx, qy, qz) < tau2;
      if (in && refit) {
        acc[0] += 1.0;
        acc[1] += px; acc[2] += py; acc[3] += pz;
        acc[4] += qx; acc[5] += qy; acc[6] += qz;
        acc[7] += qx * px; acc[8] += qx * py; acc[9] += qx * pz;
        acc[10] += qy * px; acc[11] += qy * py; acc[12] += qy * pz;
        acc[13] += qz * px; acc[14] += qz * py; acc[15] += qz * pz;
      }
    }
    if (mask) mask[k] = in ? 1 : 0;
  }
  if (refit) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      double v = acc[i];
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) red[w][i] = v;
    }
    __syncthreads();
    if (t == 0) {
      double tot[16];
      for (int i = 0; i < 16; ++i) {
        double v = 0.0;
        for (int ww = 0; ww < FIN_THREADS / 32; ++ww) v += red[ww][i];
        tot[i] = v;
      }
      const double nin = tot[0];
      if (nin >= 3.0) {
        double pm[3], qm[3], S[9], rt2[12];
        for (int i = 0; i < 3; ++i) {
          pm[i] = tot[1 + i] / nin;
          qm[i] = tot[4 + i] / nin;
        }
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) S[i * 3 + j] = tot[7 + i * 3 + j] - nin * (qm[i] * pm[j]);
        if (fit_from_sigma(S, pm, qm, rt2))
          for (int i = 0; i < 12; ++i) rt_s[i] = rt2[i];
      }
    }
    __syncthreads();
  }
  if (t < 16) {
    const int i = t >> 2, j = t & 3;
    T[t] = (i == 3) ? ((j == 3) ? 1.0 : 0.0) : ((j == 3) ? rt_s[9 + i] : rt_s[i * 3 + j]);
  }
}

// hypotheses only (no pq gather): rts[n_hyp][12], counts[h] = 0 or -1 (degenerate sample)
int ransac_hypotheses(vfmreg_ctx* ctx, const void* src_xyz, const void* tgt_xyz, int xyz_f64, const int32_t* corr, const int32_t* count,
                      int32_t max_corr, const int32_t* sample_idx, int32_t n_hyp, uint64_t seed, double* rts, int32_t* counts) {
  const int fit_blocks = ceil_div(n_hyp, PREP_THREADS);
  if (xyz_f64)
    gather_kabsch_kernel<double><<<fit_blocks, PREP_THREADS, 0, ctx->stream>>>((const double*)src_xyz, (const double*)tgt_xyz, corr, count,
                                                                              max_corr, 0, nullptr, sample_idx, n_hyp, seed, rts, counts, nullptr);
  else
    gather_kabsch_kernel<float><<<fit_blocks, PREP_THREADS, 0, ctx->stream>>>((const float*)src_xyz, (const float*)tgt_xyz, corr, count,
                                                                             max_corr, 0, nullptr, sample_idx, n_hyp, seed, rts, counts, nullptr);
  return launch_check(ctx, "gather_kabsch_kernel");
}

size_t ransac_scratch(int32_t max_corr, int32_t n_hyp) {
  return arena_bytes((size_t)(max_corr > 0 ? max_corr : 1) * 6, 8) + arena_bytes((size_t)n_hyp * 12, 8) +
         arena_bytes((size_t)n_hyp, 4) + arena_bytes((size_t)n_hyp, 8);
}

int ransac_solve(vfmreg_ctx* ctx, const void* src_xyz, const void* tgt_xyz, int xyz_f64, const int32_t* corr,
                 const int32_t* count, int32_t max_corr, const int32_t* sample_idx, int32_t n_hyp, uint64_t seed,
                 double thresh, int refit, double* T, int32_t* counts, int64_t* sumq, uint8_t* mask, int64_t* stats) {
  VFM_CHECK_ARG(n_hyp > 0, "ransac: n_hyp must be positive");
  VFM_CHECK_ARG(max_corr >= 0, "ransac: negative max_corr");
  VFM_CHECK_ARG(thresh > 0.0, "ransac: inlier threshold must be > 0 (Open3D returns an empty result otherwise)");
  const int mc = max_corr > 0 ? max_corr : 1;
  double* pq = arena_take<double>(ctx, (size_t)mc * 6);
  double* rts = arena_take<double>(ctx, (size_t)n_hyp * 12);
  int32_t* cnt = arena_take<int32_t>(ctx, (size_t)n_hyp);
  unsigned long long* sq = arena_take<unsigned long long>(ctx, (size_t)n_hyp);
  if (!pq || !rts || !cnt || !sq) {
    set_error("ransac: scratch arena too small");
    return VFMREG_ERR_ALLOC;
  }
  const double tau2 = thresh * thresh;
  const double scale = 1099511627776.0 / tau2;
  group_begin(ctx, GROUP_KABSCH);
  {
    const int gather_blocks = ceil_div(max_corr, PREP_THREADS), fit_blocks = ceil_div(n_hyp, PREP_THREADS);
    if (xyz_f64)
      gather_kabsch_kernel<double><<<gather_blocks + fit_blocks, PREP_THREADS, 0, ctx->stream>>>(
          (const double*)src_xyz, (const double*)tgt_xyz, corr, count, max_corr, gather_blocks, pq, sample_idx, n_hyp, seed, rts, cnt, sq);
    else
      gather_kabsch_kernel<float><<<gather_blocks + fit_blocks, PREP_THREADS, 0, ctx->stream>>>(
          (const float*)src_xyz, (const float*)tgt_xyz, corr, count, max_corr, gather_blocks, pq, sample_idx, n_hyp, seed, rts, cnt, sq);
    VFM_TRY(launch_check(ctx, "gather_kabsch_kernel"));
  }
  group_end(ctx, GROUP_KABSCH, 1);
  if (max_corr >= 3) {
    const int hyp_blocks = ceil_div(n_hyp, HYP_PER_CTA);
    const int max_splits = ceil_div(max_corr, 64);                  // at least 64 correspondences per slice
    static const int per_sm = [] { const char* e = getenv("VFMREG_SCORE_CTAS_PER_SM"); return e ? atoi(e) : VFM_SCORE_MIN_CTAS; }();   // tuning aid
    // per_sm CTAs of 128 threads fit an SM (registers): the grid must not exceed one wave -- rounding the slice count UP left a
    // second wave of 16 CTAs (608 CTAs on 592 slots) that doubled the kernel's time
    int splits = (int)(((int64_t)ctx->sm_count * per_sm) / hyp_blocks);
    if (splits < 1) splits = 1;
    if (splits > max_splits) splits = max_splits;
    group_begin(ctx, GROUP_RANSAC);
    score_kernel<<<dim3(hyp_blocks, splits), SCORE_THREADS, 0, ctx->stream>>>(pq, count, max_corr, rts, n_hyp, tau2, scale,
                                                                             cnt, sq);
    VFM_TRY(launch_check(ctx, "score_kernel"));
    group_end(ctx, GROUP_RANSAC, 1);
  }
  GroupScope g_fin(ctx, GROUP_FINALIZE, 1);
  finalize_kernel<<<1, FIN_THREADS, 0, ctx->stream>>>(pq, count, max_corr, rts, n_hyp, tau2, refit, cnt, sq, T, counts, sumq,
                                              mask, stats);
  return launch_check(ctx, "finalize_kernel");
}

}  // namespace vfm
