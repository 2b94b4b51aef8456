// bf16 x bf16 -> fp32 GEMM on tcgen05 for the DINOv2 forward (reference: the ViT behind image_features.py:95-101),
// C[M, N] = A[M, K] W[N, K]^T with the layer's element-wise tail fused into the epilogue so that no intermediate is
// re-read from HBM:
//   EPI_BF16_BIAS       out_bf16 = acc + bias                       (QKV projection)
//   EPI_BF16_BIAS_GELU  out_bf16 = gelu(acc + bias), exact erf      (MLP fc1)
//   EPI_F32_RESID       x += gamma * (acc + bias), fp32 in place    (attention proj / MLP fc2 + LayerScale + residual)
//   EPI_F32_PATCH       x[token row] = acc + bias + pos_embed       (patch embedding; skips the CLS row of each image)
//
// Persistent kernel, one CTA per SM, static round-robin over 128 x BN output tiles (BN = 256 when N % 256 == 0, else 192:
// every DINOv2 width is a multiple of one of them; wide N keeps the MMA off the shared-memory read limit, which a
// 128-wide tile hits).  Every weight carries a TMA box per width it divides into (256 / 192 / 128), and each
// launch picks the width with the fewest waves x tile time for its M: large batches use 256, BASELINE config 3 (6 images x
// 257 tokens = 13 row blocks) uses 192 for QKV and 128 for proj / fc1 / fc2.  Warp 0 = TMA producer (128B-swizzled 64-wide K chunks, 4-stage mbarrier ring), warp 1 = TMEM
// allocator + elected-lane tcgen05.mma issuer (kind::f16, bf16 operands, M128 N{192,256} K16), accumulators double-buffered in
// TMEM (2 x 256 columns) so the epilogue of tile i (warps 2-5: tcgen05.ld 32x32b, one output row per thread) overlaps
// the MMAs of tile i+1.  Bound: tensor pipe for the large layers; at BASELINE config 3 (6 images, 1542 tokens) most layers
// are below one wave of tiles and are latency bound.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "vit.cuh"

namespace vfm {

constexpr int GBM = 128, GBK = 64, GSTAGES = 4;
constexpr uint32_t GA_BYTES = GBM * GBK * 2;
constexpr uint32_t GB_BYTES_MAX = 256 * GBK * 2;
constexpr uint32_t GSMEM_BARS = GSTAGES * (GA_BYTES + GB_BYTES_MAX);
constexpr uint32_t GSMEM_TOTAL = GSMEM_BARS + 256 + 1024;

// exact-erf GELU.  (An Abramowitz-Stegun erf -- reciprocal + ex2 + 6 fma -- was measured SLOWER than libm's erff here:
// 17.98 vs 16.24 ms for ViT-L/14 B=48 on the same box; erff's polynomial fast path has no reciprocal.)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <int EPI, int BN>
__global__ void __launch_bounds__(192, 1)
    vit_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const GemmEpilogue ep) {
  constexpr uint32_t GB_BYTES = BN * GBK * 2;
  constexpr uint32_t IDESC = umma_idesc_f16(GBM, BN, 1);  // bf16 operands
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base, sB = base + GSTAGES * GA_BYTES;
  const uint32_t bars = base + GSMEM_BARS;
  const uint32_t full0 = bars, empty0 = bars + 8 * GSTAGES, tfull0 = bars + 16 * GSTAGES, tempty0 = tfull0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + GSMEM_BARS + 16 * GSTAGES + 32);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_count = ep.k / GBK;
  const int n_tiles = ep.n / BN;
  const int total = ((ep.m + GBM - 1) / GBM) * n_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < GSTAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    uint32_t stage = 0, phase = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const int rb = t / n_tiles, nb = t % n_tiles;   // consecutive CTAs share the activation row block (L2 reuse)
#pragma unroll 1
      for (int kb = 0; kb < kb_count; ++kb) {
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(full0 + 8 * stage, GA_BYTES + GB_BYTES);
          tma_load_2d(sA + stage * GA_BYTES, &map_a, full0 + 8 * stage, kb * GBK, rb * GBM);
          tma_load_2d(sB + stage * GB_BYTES_MAX, &map_w, full0 + 8 * stage, kb * GBK, nb * BN);
        }
        __syncwarp();
        if (++stage == GSTAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    uint32_t stage = 0, phase = 0;
    int it = 0;
    const uint64_t da0 = umma_desc_k_sw128(sA), db0 = umma_desc_k_sw128(sB);
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const uint32_t buf = (uint32_t)(it & 1);
      mbar_wait(tempty0 + 8 * buf, (uint32_t)((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * 256;
#pragma unroll 1
      for (int kb = 0; kb < kb_count; ++kb) {
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = da0 + (uint64_t)(stage * (GA_BYTES >> 4));
          const uint64_t db = db0 + (uint64_t)(stage * (GB_BYTES_MAX >> 4));
#pragma unroll
          for (int k = 0; k < GBK / 16; ++k)
            tc_mma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC, (kb | k) != 0 ? 1u : 0u);
          tc_commit(empty0 + 8 * stage);
          if (kb == kb_count - 1) tc_commit(tfull0 + 8 * buf);
        }
        __syncwarp();
        if (++stage == GSTAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int rb = t / n_tiles, nb = t % n_tiles;
      const uint32_t buf = (uint32_t)(it & 1);
      const int row = rb * GBM + q * 32 + lane;
      const bool live = row < ep.m;
      long long out_row = row;
      const float* pos_row = nullptr;
      if (EPI == EPI_F32_PATCH && live) {
        const int img = row / ep.np, p = row - img * ep.np;
        out_row = (long long)img * (ep.np + 1) + 1 + p;
        pos_row = ep.pos + (long long)(1 + p) * ep.n;
      }
      mbar_wait(tfull0 + 8 * buf, (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tc_ld32(t_addr + c * 32, r);
        tc_wait_ld();
        if (!live) continue;
        const int col0 = nb * BN + c * 32;
        if (EPI == EPI_BF16_BIAS || EPI == EPI_BF16_BIAS_GELU) {
          __nv_bfloat16* o = ep.out_bf16 + out_row * ep.ldo + col0;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint32_t pk[4];
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + i));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + i + 4));
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float v0 = __uint_as_float(r[i + 2 * j]) + bb[2 * j];
              float v1 = __uint_as_float(r[i + 2 * j + 1]) + bb[2 * j + 1];
              if (EPI == EPI_BF16_BIAS_GELU) {
                v0 = gelu_erf(v0);
                v1 = gelu_erf(v1);
              }
              __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
              pk[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(o + i) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        } else {
          float* o = ep.x + out_row * ep.ldo + col0;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            float* pv = reinterpret_cast<float*>(&v);
            if (EPI == EPI_F32_RESID) v = *reinterpret_cast<const float4*>(o + i);
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + i));
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
            float gg[4] = {0.f, 0.f, 0.f, 0.f};
            if (EPI == EPI_F32_RESID) {
              const float4 g4 = __ldg(reinterpret_cast<const float4*>(ep.gamma + col0 + i));
              gg[0] = g4.x; gg[1] = g4.y; gg[2] = g4.z; gg[3] = g4.w;
            } else {
              const float4 p4 = __ldg(reinterpret_cast<const float4*>(pos_row + col0 + i));
              gg[0] = p4.x; gg[1] = p4.y; gg[2] = p4.z; gg[3] = p4.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float a = __uint_as_float(r[i + j]) + bb[j];
              pv[j] = (EPI == EPI_F32_RESID) ? fmaf(gg[j], a, pv[j]) : a + gg[j];
            }
            *reinterpret_cast<float4*>(o + i) = v;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

template <int EPI, int BN>
static int launch_gemm(vfmreg_ctx* ctx, const CUtensorMap& a, const CUtensorMap& w, const GemmEpilogue& ep) {
  const uint64_t bit = 1ull << (EPI * 3 + (BN == 256 ? 0 : (BN == 192 ? 1 : 2)));   // per device (= per context), not per process
  if (!(ctx->gemm_attr_mask & bit)) {
    VFM_CUDA(cudaFuncSetAttribute(vit_gemm_kernel<EPI, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GSMEM_TOTAL));
    ctx->gemm_attr_mask |= bit;
  }
  const int total = ceil_div(ep.m, GBM) * (ep.n / BN);
  const int grid = total < ctx->sm_count ? total : ctx->sm_count;
  vit_gemm_kernel<EPI, BN><<<grid, 192, GSMEM_TOTAL, ctx->stream>>>(a, w, ep);
  return launch_check(ctx, "vit_gemm_kernel");
}

int vit_gemm_tile_n(int n) { return (n % 256 == 0) ? 256 : ((n % 192 == 0) ? 192 : ((n % 128 == 0) ? 128 : 0)); }

int vit_weight_maps(WeightMaps* w, const void* ptr, int n, int k) {
  const int widths[3] = {256, 192, 128};
  for (int i = 0; i < 3; ++i) {
    w->ok[i] = (n % widths[i] == 0);
    if (w->ok[i]) VFM_TRY(make_tmap_16bit(&w->map[i], ptr, n, k, k, widths[i], true));
  }
  return VFMREG_OK;
}

int vit_gemm(vfmreg_ctx* ctx, int epi, const CUtensorMap& a, const WeightMaps& w, const GemmEpilogue& ep) {
  VFM_CHECK_ARG(ep.k % GBK == 0 && ep.m > 0 && (w.ok[0] || w.ok[1] || w.ok[2]), "vit_gemm: unsupported shape m=%d n=%d k=%d", ep.m,
                ep.n, ep.k);
  // Tile width: the persistent CTAs take ceil(tiles / SMs) tiles each; a tile's MMA time is proportional to its width,
  // except that 128-wide tiles run into the shared-memory operand limit (x 1.25).  Small batches (6 images = 13 row blocks)
  // pick narrower tiles than the 256 that large batches use.  VFMREG_VIT_TILE=256|192|128 forces a width (tuning aid).
  static const int forced = [] { const char* e = getenv("VFMREG_VIT_TILE"); return e ? atoi(e) : 0; }();
  const int widths[3] = {256, 192, 128}, cost[3] = {256, 192, 160};
  int best = -1;
  long long best_cost = 0;
  for (int i = 0; i < 3; ++i) {
    if (!w.ok[i]) continue;
    const long long tiles = (long long)ceil_div(ep.m, GBM) * (ep.n / widths[i]);
    const long long c = ((tiles + ctx->sm_count - 1) / ctx->sm_count) * cost[i];
    if (forced == widths[i]) {
      best = i;
      break;
    }
    if (best < 0 || c < best_cost) {
      best = i;
      best_cost = c;
    }
  }
#define VFM_GEMM_CASE(E)                                                                              \
  case E:                                                                                             \
    return best == 0 ? launch_gemm<E, 256>(ctx, a, w.map[0], ep)                                      \
                     : (best == 1 ? launch_gemm<E, 192>(ctx, a, w.map[1], ep) : launch_gemm<E, 128>(ctx, a, w.map[2], ep));
  switch (epi) {
    VFM_GEMM_CASE(EPI_BF16_BIAS)
    VFM_GEMM_CASE(EPI_BF16_BIAS_GELU)
    VFM_GEMM_CASE(EPI_F32_RESID)
    VFM_GEMM_CASE(EPI_F32_PATCH)
  }
#undef VFM_GEMM_CASE
  set_error("vit_gemm: bad epilogue %d", epi);
  return VFMREG_ERR_ARG;
}

}  // namespace vfm
