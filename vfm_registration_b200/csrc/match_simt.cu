// Exact fp32 inner-product top-2 search (reference: faiss IndexFlatIP::search(k=1) at VoxelHashMap.cpp:486-495,
// plus the runner-up for the ratio test), CUDA-core version.
//
// score(i, j) = fmaf chain over k ascending of a[i][k] * b[j][k]  (the canonical value), so results are
// bit-identical to oracle/c/oracle_ref.c:orc_match_top2.  Ties resolve to the lowest j (faiss keeps the first
// maximum: strict '>' while scanning j ascending).
//
// Layout: a (n x dp), b (m x dp) row-major fp32 with dp % 16 == 0 (zero padded by normalize_rows).
// Tiling: CTA = 128 query rows x a span of 128-wide column tiles; 256 threads, 8x8 register tile per thread,
// K staged through shared memory in chunks of 16 with register prefetch of the next chunk (128-bit global loads,
// 128-bit conflict-free shared loads).  The per-row running (best, idx, second) lives in registers across the
// CTA's whole column span; 16 lanes share a row and are merged by shuffles at the end; the column splits are merged
// by merge_splits_kernel.  Bound: fp32 FMA pipe (arithmetic intensity n*m/(2(n+m)) FLOP/B >> ridge).
#include <float.h>

#include "common.cuh"

namespace vfm {

constexpr int BM = 128, BN = 128, BK = 16, LDS = BM + 4;

struct Top2 {
  float best;
  int idx;
  float sec;
};

// merge candidate y into x; ties on `best` go to the lower index
__device__ __forceinline__ void top2_merge(Top2& x, const Top2& y) {
  const bool y_wins = (y.best > x.best) || (y.best == x.best && y.idx < x.idx);
  const float lo = y_wins ? x.best : y.best;
  const float s = fmaxf(fmaxf(x.sec, y.sec), lo);
  if (y_wins) {
    x.best = y.best;
    x.idx = y.idx;
  }
  x.sec = s;
}

__global__ void __launch_bounds__(256)
    match_simt_kernel(const float* __restrict__ a, int n, const float* __restrict__ b, int m, int dp, int tiles_per_split,
                      float* __restrict__ pbest, int* __restrict__ pidx, float* __restrict__ psec) {
  __shared__ __align__(16) float As[2][BK][LDS];
  __shared__ __align__(16) float Bs[2][BK][LDS];
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int row0 = blockIdx.x * BM;
  const int n_col_tiles = (m + BN - 1) / BN;
  const int ct0 = blockIdx.y * tiles_per_split;
  const int ct1 = min(ct0 + tiles_per_split, n_col_tiles);
  // global->smem mapping: thread loads rows lr and lr+64, k offset lk..lk+3 of the current 16-wide K chunk
  const int lr = t >> 2, lk = (t & 3) * 4;
  const int nk = dp / BK;

  Top2 run[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) run[i] = {-INFINITY, 0x7fffffff, -INFINITY};

  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* a0 = (row0 + lr < n) ? a + (int64_t)(row0 + lr) * dp + lk : nullptr;
  const float* a1 = (row0 + lr + 64 < n) ? a + (int64_t)(row0 + lr + 64) * dp + lk : nullptr;

  for (int ct = ct0; ct < ct1; ++ct) {
    const int col0 = ct * BN;
    const float* b0 = (col0 + lr < m) ? b + (int64_t)(col0 + lr) * dp + lk : nullptr;
    const float* b1 = (col0 + lr + 64 < m) ? b + (int64_t)(col0 + lr + 64) * dp + lk : nullptr;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    float4 ra0 = a0 ? __ldg(reinterpret_cast<const float4*>(a0)) : zero4;
    float4 ra1 = a1 ? __ldg(reinterpret_cast<const float4*>(a1)) : zero4;
    float4 rb0 = b0 ? __ldg(reinterpret_cast<const float4*>(b0)) : zero4;
    float4 rb1 = b1 ? __ldg(reinterpret_cast<const float4*>(b1)) : zero4;

    for (int kc = 0; kc < nk; ++kc) {
      const int buf = kc & 1;
      As[buf][lk + 0][lr] = ra0.x; As[buf][lk + 1][lr] = ra0.y; As[buf][lk + 2][lr] = ra0.z; As[buf][lk + 3][lr] = ra0.w;
      As[buf][lk + 0][lr + 64] = ra1.x; As[buf][lk + 1][lr + 64] = ra1.y; As[buf][lk + 2][lr + 64] = ra1.z; As[buf][lk + 3][lr + 64] = ra1.w;
      Bs[buf][lk + 0][lr] = rb0.x; Bs[buf][lk + 1][lr] = rb0.y; Bs[buf][lk + 2][lr] = rb0.z; Bs[buf][lk + 3][lr] = rb0.w;
      Bs[buf][lk + 0][lr + 64] = rb1.x; Bs[buf][lk + 1][lr + 64] = rb1.y; Bs[buf][lk + 2][lr + 64] = rb1.z; Bs[buf][lk + 3][lr + 64] = rb1.w;
      __syncthreads();
      if (kc + 1 < nk) {
        const int ko = (kc + 1) * BK;
        ra0 = a0 ? __ldg(reinterpret_cast<const float4*>(a0 + ko)) : zero4;
        ra1 = a1 ? __ldg(reinterpret_cast<const float4*>(a1 + ko)) : zero4;
        rb0 = b0 ? __ldg(reinterpret_cast<const float4*>(b0 + ko)) : zero4;
        rb1 = b1 ? __ldg(reinterpret_cast<const float4*>(b1 + ko)) : zero4;
      }
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 x0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        const float4 x1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
        const float4 y0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        const float4 y1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
        const float av[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        const float bv[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      // the next iteration writes the other buffer; one barrier per chunk is enough with two buffers
    }
    __syncthreads();
    // fold this tile into the running top-2 (columns ascending inside the thread)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = col0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
        const float v = acc[i][j];
        if (col < m) {
          if (v > run[i].best) {
            run[i].sec = run[i].best;
            run[i].best = v;
            run[i].idx = col;
          } else if (v > run[i].sec) {
            run[i].sec = v;
          }
        }
      }
    }
  }
  // merge the 16 lanes (tx) that share each row; they are 16 consecutive lanes of a warp
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
      Top2 o;
      o.best = __shfl_xor_sync(0xffffffffu, run[i].best, off);
      o.idx = __shfl_xor_sync(0xffffffffu, run[i].idx, off);
      o.sec = __shfl_xor_sync(0xffffffffu, run[i].sec, off);
      top2_merge(run[i], o);
    }
    if (tx == 0) {
      const int row = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
      if (row < n) {
        const int64_t o = (int64_t)blockIdx.y * n + row;
        pbest[o] = run[i].best;
        pidx[o] = run[i].idx;
        psec[o] = run[i].sec;
      }
    }
  }
}

__global__ void merge_splits_kernel(const float* __restrict__ pbest, const int* __restrict__ pidx,
                                    const float* __restrict__ psec, int n, int splits, int32_t* __restrict__ idx,
                                    float* __restrict__ best, float* __restrict__ sec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Top2 r = {pbest[i], pidx[i], psec[i]};
  for (int s = 1; s < splits; ++s) {
    const Top2 o = {pbest[(int64_t)s * n + i], pidx[(int64_t)s * n + i], psec[(int64_t)s * n + i]};
    top2_merge(r, o);
  }
  idx[i] = (r.idx == 0x7fffffff) ? 0 : r.idx;
  if (best) best[i] = r.best;
  if (sec) sec[i] = r.sec;
}

static int pick_splits(vfmreg_ctx* ctx, int64_t n, int64_t m, int* tiles_per_split) {
  const int row_blocks = ceil_div(n, BM);
  const int col_tiles = ceil_div(m, BN);
  // aim for >= 4 CTAs per SM (2 resident) so the tail wave is small
  int want = ceil_div((int64_t)ctx->sm_count * 4, row_blocks);
  if (want < 1) want = 1;
  if (want > col_tiles) want = col_tiles;
  *tiles_per_split = ceil_div(col_tiles, want);
  return ceil_div(col_tiles, *tiles_per_split);
}

size_t match_simt_scratch(vfmreg_ctx* ctx, int64_t n, int64_t m) {
  int tps;
  const int splits = pick_splits(ctx, n, m, &tps);
  return 3 * arena_bytes((size_t)splits * n, 4);
}

int match_simt(vfmreg_ctx* ctx, const float* a, int64_t n, const float* b, int64_t m, int dp, int32_t* idx,
               float* best, float* sec) {
  VFM_CHECK_ARG(n > 0 && m > 0 && n < (1LL << 31) && m < (1LL << 31), "match_simt: bad sizes n=%lld m=%lld",
                (long long)n, (long long)m);
  VFM_CHECK_ARG(dp % BK == 0, "match_simt: padded dim %d not a multiple of %d", dp, BK);
  int tps;
  const int splits = pick_splits(ctx, n, m, &tps);
  float* pbest = arena_take<float>(ctx, (size_t)splits * n);
  int* pidx = arena_take<int>(ctx, (size_t)splits * n);
  float* psec = arena_take<float>(ctx, (size_t)splits * n);
  if (!pbest || !pidx || !psec) {
    set_error("match_simt: scratch arena too small");
    return VFMREG_ERR_ALLOC;
  }
  dim3 grid(ceil_div(n, BM), splits);
  group_begin(ctx, GROUP_MATCH);
  match_simt_kernel<<<grid, 256, 0, ctx->stream>>>(a, (int)n, b, (int)m, dp, tps, pbest, pidx, psec);
  VFM_TRY(launch_check(ctx, "match_simt_kernel"));
  group_end(ctx, GROUP_MATCH, 1);
  merge_splits_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(pbest, pidx, psec, (int)n, splits, idx, best, sec);
  return launch_check(ctx, "merge_splits_kernel");
}

}  // namespace vfm
