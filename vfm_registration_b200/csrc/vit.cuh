// Internal declarations of the DINOv2 forward (vit.cu / vit_gemm.cu / vit_ops.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace vfm {

enum { EPI_BF16_BIAS = 0, EPI_BF16_BIAS_GELU = 1, EPI_F32_RESID = 2, EPI_F32_PATCH = 3 };

struct GemmEpilogue {
  int m, n, k;              // problem size (k already padded to a multiple of 64)
  int np;                   // EPI_F32_PATCH: patches per image
  long long ldo;            // row pitch of the output in elements
  const float* bias;        // [n]
  const float* gamma;       // [n]  LayerScale (EPI_F32_RESID)
  const float* pos;         // [(1 + np), n] position embedding (EPI_F32_PATCH)
  float* x;                 // fp32 residual stream (EPI_F32_*)
  __nv_bfloat16* out_bf16;  // bf16 output (EPI_BF16_*)
};

// One weight matrix (N x K, K-major bf16) with a TMA box per output-tile width it can be tiled by (256 / 192 / 128 rows).
struct WeightMaps {
  CUtensorMap map[3];    // [0] = 256-row box, [1] = 192, [2] = 128
  bool ok[3] = {false, false, false};
};
int vit_weight_maps(WeightMaps* w, const void* ptr, int n, int k);
// picks the tile width that needs the fewest (waves x tile time) for this M on this GPU
int vit_gemm(vfmreg_ctx* ctx, int epi, const CUtensorMap& a, const WeightMaps& w, const GemmEpilogue& ep);
// output-tile width the GEMM uses for an N-column weight (256 or 192; 0 = unsupported): the weight's TMA box has that many rows
int vit_gemm_tile_n(int n);

// uint8 HWC images -> resized, normalised, im2col'ed bf16 patch matrix (B*np x kp) + CLS rows of the residual stream
int vit_preprocess(vfmreg_ctx* ctx, const uint8_t* images, int b, int h, int w, int gh, int gw, int patch, const float* mean_std,
                   __nv_bfloat16* patches, int kp, float* x, const float* cls, const float* pos, int width);
int vit_layernorm_bf16(vfmreg_ctx* ctx, const float* x, int rows, int width, const float* g, const float* b, float eps,
                       __nv_bfloat16* out);
int vit_attention(vfmreg_ctx* ctx, const __nv_bfloat16* qkv, int b, int t, int heads, int width, __nv_bfloat16* out);
int vit_final_norm(vfmreg_ctx* ctx, const float* x, int b, int t, int width, const float* g1, const float* b1, float eps1,
                   const float* g2, const float* b2, float eps2, int channel_norm, float* out);

}  // namespace vfm
