// Internal declarations of the DINOv2 forward (vit.cu / vit_gemm.cu / vit_ops.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace vfm {

enum { EPI_BF16_BIAS = 0, EPI_BF16_BIAS_GELU = 1, EPI_F32_PARTIAL = 2, EPI_F32_PATCH = 3, EPI_F32_RESID = 4 };

struct GemmEpilogue {
  int m, n, k;              // tokens (rows of X), output features (rows of W), k already padded to a multiple of 64
  int np;                   // EPI_F32_PATCH: patches per image
  int nt, n_full, tail_w, fb_count, split;   // filled in by vit_gemm from its plan: token tile width, full token tiles, width of
                                             // the tail tile (0 = none, else a multiple of 16), 256-feature blocks, K splits
  long long ldo;            // row pitch of the output in elements
  const float* bias;        // [n]
  const float* scale;       // [n] LayerScale gamma (EPI_F32_RESID)
  const float* pos;         // [(1 + np), n] position embedding (EPI_F32_PATCH)
  float* x;                 // EPI_F32_PATCH / EPI_F32_RESID: fp32 residual stream; EPI_F32_PARTIAL: workspace [split][m][ldo]
  __nv_bfloat16* out_bf16;  // bf16 output (EPI_BF16_*)
  const char* pf_ptr;       // optional: byte range prefetched into L2 while this GEMM runs (the weights of a later GEMM)
  unsigned long long pf_bytes;
};

// TMA map of one weight matrix (N x K, K-major bf16): box = 128 rows (one CTA's features) x 64 k
int vit_weight_map(CUtensorMap* map, const void* ptr, int n, int k);
// How one GEMM of the forward is cut into units: token tile width (multiple of 32, <= 256), K splits (EPI_F32_PARTIAL only),
// the tail tile and the unit count; picked by a cost model for m tokens x n features x k on this GPU (vit_gemm.cu).
struct GemmPlan {
  int nt, split, tail_w, n_full, fb_count, units;
};
GemmPlan vit_gemm_plan(vfmreg_ctx* ctx, int m, int n, int k, int max_split, const char* which);
// TMA maps of a token matrix (rows x k, K-major bf16) for a plan: boxes of nt / 2 rows and of tail_w / 2 rows
struct TokenMaps {
  CUtensorMap full, tail;
};
int vit_token_maps(TokenMaps* t, const void* ptr, int rows, int k, const GemmPlan& plan);
// out[token, feature] = X W^T with the epilogue `epi`
int vit_gemm(vfmreg_ctx* ctx, int epi, const CUtensorMap& w, const TokenMaps& x, const GemmPlan& plan, GemmEpilogue ep);
// launch helper: programmatic stream serialisation (the kernel must call pdl_wait() before its first global access)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// uint8 HWC images -> resized, normalised, im2col'ed bf16 patch matrix (B*np x kp) + CLS rows of the residual stream
int vit_preprocess(vfmreg_ctx* ctx, const uint8_t* images, int b, int h, int w, int gh, int gw, int patch, const float* mean_std,
                   __nv_bfloat16* patches, int kp, float* x, const float* cls, const float* pos, int width);
// Pending residual branch of the previous GEMM (EPI_F32_PARTIAL): x += scale * (sum over splits of ws[s] + bias), applied by
// the normalisation kernel that follows the GEMM (splits summed in ascending order: deterministic)
struct Residual {
  const float* ws = nullptr;   // [split][rows][width], null = nothing pending
  int split = 0;
  const float* bias = nullptr;
  const float* scale = nullptr;   // LayerScale gamma
};
// x (+= pending residual, written back) -> LayerNorm -> bf16
int vit_layernorm_bf16(vfmreg_ctx* ctx, float* x, int rows, int width, const Residual& res, const float* g, const float* b, float eps,
                       __nv_bfloat16* out);
// tcgen05 attention (vit_attn.cu) for t <= 320 tokens per image; map_qkv: the (rows, 3 width) QKV matrix with 64 x 64 boxes
bool vit_attention_tc_supported(int t);
// pf_ptr / pf_bytes: a byte range (the next GEMM's weights) the kernel prefetches into L2 while it runs (may be null)
int vit_attention_tc(vfmreg_ctx* ctx, const CUtensorMap& map_qkv, int b, int t, int heads, int width, __nv_bfloat16* out,
                     const void* pf_ptr, size_t pf_bytes);
int vit_attention(vfmreg_ctx* ctx, const __nv_bfloat16* qkv, int b, int t, int heads, int width, __nv_bfloat16* out);
int vit_final_norm(vfmreg_ctx* ctx, const float* x, int b, int t, int width, const Residual& res, const float* g1, const float* b1, float eps1,
                   const float* g2, const float* b2, float eps2, int channel_norm, float* out);

}  // namespace vfm
