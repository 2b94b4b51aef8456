// bf16 x bf16 -> fp32 GEMM on tcgen05 for the DINOv2 forward (reference: the ViT behind image_features.py:95-101),
// C[M, N] = A[M, K] W[N, K]^T with the layer's element-wise tail fused into the epilogue so that no intermediate is
// re-read from HBM:
//   EPI_BF16_BIAS       out_bf16 = acc + bias                       (QKV projection)
//   EPI_BF16_BIAS_GELU  out_bf16 = gelu(acc + bias), exact erf      (MLP fc1)
//   EPI_F32_RESID       x += gamma * (acc + bias), fp32 in place    (attention proj / MLP fc2 + LayerScale + residual)
//   EPI_F32_PATCH       x[token row] = acc + bias + pos_embed       (patch embedding; skips the CLS row of each image)
//
// One CTA per 128 x BN output tile (BN = 128): warp 0 = TMA producer (128B-swizzled 64-wide K chunks, 4-stage mbarrier
// ring), warp 1 = TMEM allocator + single-lane tcgen05.mma issuer (kind::f16, bf16 operands, M128 N128 K16), warps 2-5 =
// epilogue (tcgen05.ld 32x32b, one output row per thread).  Bound: tensor pipe for the large layers; at BASELINE
// config 3 (6 images, 1542 tokens) the grids are below one wave and the layer is launch/latency bound.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "vit.cuh"

namespace vfm {

constexpr int GBM = 128, GBN = 128, GBK = 64, GSTAGES = 4;
constexpr uint32_t GA_BYTES = GBM * GBK * 2, GB_BYTES = GBN * GBK * 2;
constexpr uint32_t GSMEM_BARS = GSTAGES * (GA_BYTES + GB_BYTES);
constexpr uint32_t GSMEM_TOTAL = GSMEM_BARS + 128 + 1024;
constexpr uint32_t G_IDESC = umma_idesc_f16(GBM, GBN, 1);  // bf16 operands

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <int EPI>
__global__ void __launch_bounds__(192, 1)
    vit_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const GemmEpilogue ep) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base, sB = base + GSTAGES * GA_BYTES;
  const uint32_t bars = base + GSMEM_BARS;
  const uint32_t full0 = bars, empty0 = bars + 8 * GSTAGES, tfull = bars + 16 * GSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + GSMEM_BARS + 16 * GSTAGES + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rb = blockIdx.x, nb = blockIdx.y;
  const int kb_count = ep.k / GBK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < GSTAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), GBN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    uint32_t stage = 0, phase = 0;
#pragma unroll 1
    for (int kb = 0; kb < kb_count; ++kb) {
      mbar_wait(empty0 + 8 * stage, phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(full0 + 8 * stage, GA_BYTES + GB_BYTES);
        tma_load_2d(sA + stage * GA_BYTES, &map_a, full0 + 8 * stage, kb * GBK, rb * GBM);
        tma_load_2d(sB + stage * GB_BYTES, &map_w, full0 + 8 * stage, kb * GBK, nb * GBN);
      }
      __syncwarp();
      if (++stage == GSTAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    uint32_t stage = 0, phase = 0;
    const uint64_t da0 = umma_desc_k_sw128(sA), db0 = umma_desc_k_sw128(sB);
#pragma unroll 1
    for (int kb = 0; kb < kb_count; ++kb) {
      mbar_wait(full0 + 8 * stage, phase);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = da0 + (uint64_t)(stage * (GA_BYTES >> 4));
        const uint64_t db = db0 + (uint64_t)(stage * (GB_BYTES >> 4));
#pragma unroll
        for (int k = 0; k < GBK / 16; ++k)
          tc_mma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), G_IDESC, (kb | k) != 0 ? 1u : 0u);
        tc_commit(empty0 + 8 * stage);
        if (kb == kb_count - 1) tc_commit(tfull);
      }
      __syncwarp();
      if (++stage == GSTAGES) { stage = 0; phase ^= 1; }
    }
  } else {
    const int q = warp & 3;
    const int row = rb * GBM + q * 32 + lane;
    mbar_wait(tfull, 0);
    tc_fence_after();
    const bool live = row < ep.m;
    long long out_row = row;
    const float* pos_row = nullptr;
    if (EPI == EPI_F32_PATCH && live) {
      const int img = row / ep.np, p = row - img * ep.np;
      out_row = (long long)img * (ep.np + 1) + 1 + p;
      pos_row = ep.pos + (long long)(1 + p) * ep.n;
    }
#pragma unroll 1
    for (int c = 0; c < GBN / 32; ++c) {
      uint32_t r[32];
      tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, r);
      tc_wait_ld();
      if (!live) continue;
      const int col0 = nb * GBN + c * 32;
      if (EPI == EPI_BF16_BIAS || EPI == EPI_BF16_BIAS_GELU) {
        __nv_bfloat16* o = ep.out_bf16 + out_row * ep.ldo + col0;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float v0 = __uint_as_float(r[i + 2 * j]) + __ldg(ep.bias + col0 + i + 2 * j);
            float v1 = __uint_as_float(r[i + 2 * j + 1]) + __ldg(ep.bias + col0 + i + 2 * j + 1);
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
          float4 v;
          float* pv = reinterpret_cast<float*>(&v);
          if (EPI == EPI_F32_RESID) v = *reinterpret_cast<const float4*>(o + i);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float a = __uint_as_float(r[i + j]) + __ldg(ep.bias + col0 + i + j);
            if (EPI == EPI_F32_RESID)
              pv[j] = fmaf(__ldg(ep.gamma + col0 + i + j), a, pv[j]);
            else
              pv[j] = a + __ldg(pos_row + col0 + i + j);
          }
          *reinterpret_cast<float4*>(o + i) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, GBN);
  }
}

template <int EPI>
static int launch_gemm(vfmreg_ctx* ctx, const CUtensorMap& a, const CUtensorMap& w, const GemmEpilogue& ep) {
  static bool attr_set = false;
  if (!attr_set) {
    VFM_CUDA(cudaFuncSetAttribute(vit_gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GSMEM_TOTAL));
    attr_set = true;
  }
  dim3 grid(ceil_div(ep.m, GBM), ep.n / GBN);
  vit_gemm_kernel<EPI><<<grid, 192, GSMEM_TOTAL, ctx->stream>>>(a, w, ep);
  return launch_check(ctx, "vit_gemm_kernel");
}

int vit_gemm(vfmreg_ctx* ctx, int epi, const CUtensorMap& a, const CUtensorMap& w, const GemmEpilogue& ep) {
  VFM_CHECK_ARG(ep.k % GBK == 0 && ep.n % GBN == 0 && ep.m > 0, "vit_gemm: unsupported shape m=%d n=%d k=%d", ep.m, ep.n, ep.k);
  switch (epi) {
    case EPI_BF16_BIAS: return launch_gemm<EPI_BF16_BIAS>(ctx, a, w, ep);
    case EPI_BF16_BIAS_GELU: return launch_gemm<EPI_BF16_BIAS_GELU>(ctx, a, w, ep);
    case EPI_F32_RESID: return launch_gemm<EPI_F32_RESID>(ctx, a, w, ep);
    case EPI_F32_PATCH: return launch_gemm<EPI_F32_PATCH>(ctx, a, w, ep);
  }
  set_error("vit_gemm: bad epilogue %d", epi);
  return VFMREG_ERR_ARG;
}

}  // namespace vfm
