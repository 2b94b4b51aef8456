// Non-GEMM kernels of the DINOv2 forward (reference: image_features.py:67-77 preprocessing, :95-101 model call; the
// network is third party -- architecture per SURVEY.md A.9).
//   preprocess_kernel   ToTensor -> bilinear Resize(antialias=False, align_corners=False) -> Normalize, written straight
//                       into the im2col patch matrix in bf16 (the resized image is never materialised); HBM-bound.
//   layernorm kernels   one warp per token, 128-bit loads, two-pass statistics in registers; HBM-bound.
//   attention_kernel    softmax(Q K^T / sqrt(64)) V, one CTA per (image, head, query split): K/V of the head staged in shared
//                       memory, 8 warps walk 16-query tiles with a flash-style online softmax over 64-key chunks, bf16
//                       mma.sync m16n8k16 with fp32 accumulation.  (~4 % of the model's FLOPs; the GEMMs that carry
//                       the rest run on tcgen05.)
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "vit.cuh"

namespace vfm {

// ---------------------------------------------------------------------------------------------------------------------
// PATCH > 0: the patch size as a compile-time constant (14 for every DINOv2 model: the six divisions by it and by its square
// become multiplications), 0: the run-time value
template <int PATCH>
__global__ void __launch_bounds__(256)
    preprocess_kernel(const uint8_t* __restrict__ images, int b, int h, int w, int gh, int gw, int patch_rt, float m0, float m1,
                      float m2, float s0, float s1, float s2, __nv_bfloat16* __restrict__ patches, int kp) {
  const int patch = PATCH > 0 ? PATCH : patch_rt;
  pdl_launch_dependents();
  const long long total = (long long)b * gh * gw * kp;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  const int k = (int)(g % kp);
  const long long pr = g / kp;
  const int pp = patch * patch;
  if (k >= 3 * pp) {
    patches[g] = __float2bfloat16(0.f);
    return;
  }
  const int np = gh * gw;
  const int img = (int)(pr / np), p = (int)(pr % np);
  const int py = p / gw, px = p % gw;
  const int c = k / pp, rem = k % pp;
  const int oy = py * patch + rem / patch, ox = px * patch + rem % patch;
  const int out_h = gh * patch, out_w = gw * patch;
  const float sy = (float)h / (float)out_h, sx = (float)w / (float)out_w;
  float fy = sy * ((float)oy + 0.5f) - 0.5f, fx = sx * ((float)ox + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy;
  fx = fx < 0.f ? 0.f : fx;
  const int y0 = min((int)fy, h - 1), x0 = min((int)fx, w - 1);
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
  const uint8_t* im = images + (long long)img * h * w * 3;
  const float p00 = (float)im[((long long)y0 * w + x0) * 3 + c] / 255.0f, p01 = (float)im[((long long)y0 * w + x1) * 3 + c] / 255.0f;
  const float p10 = (float)im[((long long)y1 * w + x0) * 3 + c] / 255.0f, p11 = (float)im[((long long)y1 * w + x1) * 3 + c] / 255.0f;
  const float v = hy * (hx * p00 + lx * p01) + ly * (hx * p10 + lx * p11);
  const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
  patches[g] = __float2bfloat16((v - mean) / sd);
}

__global__ void cls_rows_kernel(float* __restrict__ x, int b, int t, int width, const float* __restrict__ cls,
                                const float* __restrict__ pos) {
  pdl_launch_dependents();
  pdl_wait();   // keeps the chain of programmatic dependencies transitive (this kernel itself depends on nothing)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b * width) return;
  const int img = i / width, c = i % width;
  x[(long long)img * t * width + c] = cls[c] + pos[c];
}

int vit_preprocess(vfmreg_ctx* ctx, const uint8_t* images, int b, int h, int w, int gh, int gw, int patch, const float* ms,
                   __nv_bfloat16* patches, int kp, float* x, const float* cls, const float* pos, int width) {
  const long long total = (long long)b * gh * gw * kp;
  if (patch == 14)
    preprocess_kernel<14><<<ceil_div(total, 256), 256, 0, ctx->stream>>>(images, b, h, w, gh, gw, patch, ms[0], ms[1], ms[2], ms[3],
                                                                        ms[4], ms[5], patches, kp);
  else
    preprocess_kernel<0><<<ceil_div(total, 256), 256, 0, ctx->stream>>>(images, b, h, w, gh, gw, patch, ms[0], ms[1], ms[2], ms[3],
                                                                       ms[4], ms[5], patches, kp);
  VFM_TRY(launch_check(ctx, "preprocess_kernel"));
  VFM_CUDA(launch_pdl(cls_rows_kernel, dim3(ceil_div((long long)b * width, 256)), dim3(256), 0, ctx->stream, x, b, gh * gw + 1, width,
                      cls, pos));
  return launch_check(ctx, "cls_rows_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
constexpr int LN_MAX_V4 = 8;  // width <= 1024

struct RowStats {
  float mean, rstd;
};

__device__ __forceinline__ RowStats warp_row_stats(const float4* v, int nv, int width, float eps) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i)
    if (i < nv) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float mean = s / (float)width;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i)
    if (i < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
  return {mean, rsqrtf(q / (float)width + eps)};
}

// v += scale * (ws[0] + ws[1] + ... + bias) for one row held by a warp (the pending residual branch of the previous GEMM).
// SPLIT is a template parameter so that every load of the row is in flight before the first add (a run-time loop over the
// splits serialised 8 x SPLIT dependent L2 round trips per row: 8 us per LayerNorm at 6 images).
template <int SPLIT>
__device__ __forceinline__ void add_residual_n(float4* v, int nv, const Residual& res, long long row, int rows, int width, int lane) {
  float4 a[LN_MAX_V4];
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i)
    if (i < nv) a[i] = *reinterpret_cast<const float4*>(res.ws + row * width + i * 128 + lane * 4);
#pragma unroll
  for (int s = 1; s < SPLIT; ++s) {
    float4 p[LN_MAX_V4];
#pragma unroll
    for (int i = 0; i < LN_MAX_V4; ++i)
      if (i < nv) p[i] = *reinterpret_cast<const float4*>(res.ws + ((long long)s * rows + row) * width + i * 128 + lane * 4);
#pragma unroll
    for (int i = 0; i < LN_MAX_V4; ++i)
      if (i < nv) {
        a[i].x += p[i].x; a[i].y += p[i].y; a[i].z += p[i].z; a[i].w += p[i].w;
      }
  }
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i)
    if (i < nv) {
      const int k = i * 128 + lane * 4;
      const float4 bb = __ldg(reinterpret_cast<const float4*>(res.bias + k)), sc = __ldg(reinterpret_cast<const float4*>(res.scale + k));
      v[i].x = fmaf(sc.x, a[i].x + bb.x, v[i].x);
      v[i].y = fmaf(sc.y, a[i].y + bb.y, v[i].y);
      v[i].z = fmaf(sc.z, a[i].z + bb.z, v[i].z);
      v[i].w = fmaf(sc.w, a[i].w + bb.w, v[i].w);
    }
}
__device__ __forceinline__ void add_residual(float4* v, int nv, const Residual& res, long long row, int rows, int width, int lane) {
  switch (res.split) {   // warp-uniform
    case 1: add_residual_n<1>(v, nv, res, row, rows, width, lane); break;
    case 2: add_residual_n<2>(v, nv, res, row, rows, width, lane); break;
    case 3: add_residual_n<3>(v, nv, res, row, rows, width, lane); break;
    default: add_residual_n<4>(v, nv, res, row, rows, width, lane); break;
  }
}

// HAS_RES = false (nothing pending: the GEMM before added its branch into x itself, EPI_F32_RESID) holds one row copy instead
// of two: 4 CTAs per SM instead of 2 keep twice the bytes in flight -- this kernel is pure HBM / L2 traffic.
template <bool HAS_RES>
__global__ void __launch_bounds__(256, HAS_RES ? 2 : 4)
    layernorm_bf16_kernel(float* __restrict__ x, int rows, int width, const Residual res, const float* __restrict__ g,
                          const float* __restrict__ b, float eps, __nv_bfloat16* __restrict__ out) {
  TraceScope trace(10);
  pdl_launch_dependents();
  pdl_wait();
  trace.waited();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = width / 128;
  float* xr = x + (long long)row * width;
  float4 v[LN_MAX_V4];
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i)
    if (i < nv) v[i] = *reinterpret_cast<const float4*>(xr + i * 128 + lane * 4);
  if (HAS_RES) {
    add_residual(v, nv, res, row, rows, width, lane);
#pragma unroll
    for (int i = 0; i < LN_MAX_V4; ++i)
      if (i < nv) *reinterpret_cast<float4*>(xr + i * 128 + lane * 4) = v[i];
  }
  const RowStats st = warp_row_stats(v, nv, width, eps);
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i)
    if (i < nv) {
      const int k = i * 128 + lane * 4;
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g + k)), bb = __ldg(reinterpret_cast<const float4*>(b + k));
      __nv_bfloat162 h0 = __floats2bfloat162_rn((v[i].x - st.mean) * st.rstd * gg.x + bb.x, (v[i].y - st.mean) * st.rstd * gg.y + bb.y);
      __nv_bfloat162 h1 = __floats2bfloat162_rn((v[i].z - st.mean) * st.rstd * gg.z + bb.z, (v[i].w - st.mean) * st.rstd * gg.w + bb.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&h0);
      pk.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(out + (long long)row * width + k) = pk;
    }
  trace.end();
}

int vit_layernorm_bf16(vfmreg_ctx* ctx, float* x, int rows, int width, const Residual& res, const float* g, const float* b, float eps,
                       __nv_bfloat16* out) {
  VFM_CHECK_ARG(width % 128 == 0 && width <= 128 * LN_MAX_V4, "layernorm: width %d unsupported", width);
  if (res.ws)
    VFM_CUDA(launch_pdl(layernorm_bf16_kernel<true>, dim3(ceil_div(rows, 8)), dim3(256), 0, ctx->stream, x, rows, width, res, g, b, eps, out));
  else
    VFM_CUDA(launch_pdl(layernorm_bf16_kernel<false>, dim3(ceil_div(rows, 8)), dim3(256), 0, ctx->stream, x, rows, width, res, g, b, eps, out));
  return launch_check(ctx, "layernorm_bf16_kernel");
}

// final LayerNorm -> drop CLS -> ChannelNorm (LayerNorm over C), fp32 out (B, gh*gw, C)
__global__ void __launch_bounds__(256)
    final_norm_kernel(const float* __restrict__ x, int b, int t, int width, const Residual res, const float* __restrict__ g1, const float* __restrict__ b1,
                      float eps1, const float* __restrict__ g2, const float* __restrict__ b2, float eps2, int channel_norm,
                      float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int np = t - 1;
  if (tok >= b * np) return;
  const int img = tok / np, p = tok % np;
  const int nv = width / 128;
  const long long row = (long long)img * t + 1 + p;
  const float* xr = x + row * width;
  float4 v[LN_MAX_V4];
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i)
    if (i < nv) v[i] = *reinterpret_cast<const float4*>(xr + i * 128 + lane * 4);
  if (res.ws) add_residual(v, nv, res, row, b * t, width, lane);   // the last layer's MLP branch
  RowStats st = warp_row_stats(v, nv, width, eps1);
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i)
    if (i < nv) {
      const int k = i * 128 + lane * 4;
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g1 + k)), bb = __ldg(reinterpret_cast<const float4*>(b1 + k));
      v[i].x = (v[i].x - st.mean) * st.rstd * gg.x + bb.x;
      v[i].y = (v[i].y - st.mean) * st.rstd * gg.y + bb.y;
      v[i].z = (v[i].z - st.mean) * st.rstd * gg.z + bb.z;
      v[i].w = (v[i].w - st.mean) * st.rstd * gg.w + bb.w;
    }
  if (channel_norm) {
    st = warp_row_stats(v, nv, width, eps2);
#pragma unroll
    for (int i = 0; i < LN_MAX_V4; ++i)
      if (i < nv) {
        const int k = i * 128 + lane * 4;
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g2 + k)), bb = __ldg(reinterpret_cast<const float4*>(b2 + k));
        v[i].x = (v[i].x - st.mean) * st.rstd * gg.x + bb.x;
        v[i].y = (v[i].y - st.mean) * st.rstd * gg.y + bb.y;
        v[i].z = (v[i].z - st.mean) * st.rstd * gg.z + bb.z;
        v[i].w = (v[i].w - st.mean) * st.rstd * gg.w + bb.w;
      }
  }
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i)
    if (i < nv) *reinterpret_cast<float4*>(out + (long long)tok * width + i * 128 + lane * 4) = v[i];
}

int vit_final_norm(vfmreg_ctx* ctx, const float* x, int b, int t, int width, const Residual& res, const float* g1, const float* b1, float eps1,
                   const float* g2, const float* b2, float eps2, int channel_norm, float* out) {
  VFM_CHECK_ARG(width % 128 == 0 && width <= 128 * LN_MAX_V4, "final_norm: width %d unsupported", width);
  VFM_CUDA(launch_pdl(final_norm_kernel, dim3(ceil_div((long long)b * (t - 1), 8)), dim3(256), 0, ctx->stream, x, b, t, width, res, g1, b1,
                      eps1, g2, b2, eps2, channel_norm, out));
  return launch_check(ctx, "final_norm_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
constexpr int ATT_DH = 64, ATT_LD = 72;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

constexpr int ATT_WARPS = 8;

// One CTA per (image, head, query split): K and V of the head are staged in shared memory once per CTA, then each of the
// 8 warps walks the split's 16-query tiles (tile = first + warp, first + warp + 8, ...) with a flash-style online softmax
// over 64-key chunks.  gridDim.z query splits keep small batches from leaving SMs idle (6 images x 16 heads = 96 CTAs).
__global__ void __launch_bounds__(ATT_WARPS * 32)
    attention_kernel(const __nv_bfloat16* __restrict__ qkv, int t, int width, int tp, __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  __nv_bfloat16* ks = reinterpret_cast<__nv_bfloat16*>(att_smem);
  __nv_bfloat16* vs = ks + (size_t)tp * ATT_LD;
  __nv_bfloat16* qs_all = vs + (size_t)tp * ATT_LD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int head = blockIdx.x, img = blockIdx.y;
  const long long ld = 3LL * width;
  const __nv_bfloat16* base = qkv + (long long)img * t * ld + head * ATT_DH;
  const uint4 z4 = make_uint4(0, 0, 0, 0);
  TraceScope trace(11);
  pdl_launch_dependents();
  pdl_wait();
  trace.waited();
  for (int i = tid; i < tp * 8; i += ATT_WARPS * 32) {
    const int r = i >> 3, ch = i & 7;
    uint4 kv = z4, vv = z4;
    if (r < t) {
      kv = *reinterpret_cast<const uint4*>(base + r * ld + width + ch * 8);
      vv = *reinterpret_cast<const uint4*>(base + r * ld + 2 * width + ch * 8);
    }
    *reinterpret_cast<uint4*>(ks + r * ATT_LD + ch * 8) = kv;
    *reinterpret_cast<uint4*>(vs + r * ATT_LD + ch * 8) = vv;
  }
  __syncthreads();
  __nv_bfloat16* qs = qs_all + warp * 16 * ATT_LD;   // this warp's private 16 x 64 query tile
  const int g = lane >> 2, tq = lane & 3;
  const float sl2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  __nv_bfloat16* ob = out + (long long)img * t * width + head * ATT_DH;
  const int n_tiles = (t + 15) / 16, per_split = (n_tiles + gridDim.z - 1) / gridDim.z;
  const int q_begin = blockIdx.z * per_split * 16, q_end = min(t, (int)(blockIdx.z + 1) * per_split * 16);
  for (int q0 = q_begin + warp * 16; q0 < q_end; q0 += ATT_WARPS * 16) {
    __syncwarp();
    for (int i = lane; i < 16 * 8; i += 32) {
      const int r = i >> 3, ch = i & 7;
      uint4 qv = z4;
      if (q0 + r < t) qv = *reinterpret_cast<const uint4*>(base + (long long)(q0 + r) * ld + ch * 8);
      *reinterpret_cast<uint4*>(qs + r * ATT_LD + ch * 8) = qv;
    }
    __syncwarp();
    uint32_t qa[4][4];
    {
      const int row = (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t addr = (uint32_t)__cvta_generic_to_shared(qs + row * ATT_LD + kk * 16 + 8 * (lane >> 4));
        ldsm_x4(addr, qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3]);
      }
    }
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    for (int kc = 0; kc < tp / 64; ++kc) {
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int np2 = 0; np2 < 4; ++np2) {  // two key tiles (16 keys) per ldmatrix.x4
          const int key = kc * 64 + np2 * 16 + (lane & 7) + 8 * (lane >> 4);
          const int col = kk * 16 + 8 * ((lane >> 3) & 1);
          uint32_t b0, b1, b2, b3;
          ldsm_x4((uint32_t)__cvta_generic_to_shared(ks + key * ATT_LD + col), b0, b1, b2, b3);
          mma_bf16(s[np2 * 2], qa[kk], b0, b1);
          mma_bf16(s[np2 * 2 + 1], qa[kk], b2, b3);
        }
      }
      // mask keys beyond t, running max
      float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = kc * 64 + nt * 8 + tq * 2;
        if (key >= t) s[nt][0] = s[nt][2] = -INFINITY;
        if (key + 1 >= t) s[nt][1] = s[nt][3] = -INFINITY;
        mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
        mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      }
      float alpha[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        alpha[r] = exp2f((m_run[r] - mx[r]) * sl2);  // first chunk: exp2(-inf) = 0
        m_run[r] = mx[r];
        l_run[r] *= alpha[r];
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        o[nt][0] *= alpha[0];
        o[nt][1] *= alpha[0];
        o[nt][2] *= alpha[1];
        o[nt][3] *= alpha[1];
      }
      uint32_t pa[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float p0 = exp2f((s[nt][0] - mx[0]) * sl2), p1 = exp2f((s[nt][1] - mx[0]) * sl2);
        const float p2 = exp2f((s[nt][2] - mx[1]) * sl2), p3 = exp2f((s[nt][3] - mx[1]) * sl2);
        l_run[0] += p0 + p1;
        l_run[1] += p2 + p3;
        pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {      // 16 keys per step
#pragma unroll
        for (int dp2 = 0; dp2 < 4; ++dp2) {  // two dh tiles (16 channels) per ldmatrix.x4.trans
          const int key = kc * 64 + j * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
          const int col = dp2 * 16 + 8 * (lane >> 4);
          uint32_t b0, b1, b2, b3;
          ldsm_x4_trans((uint32_t)__cvta_generic_to_shared(vs + key * ATT_LD + col), b0, b1, b2, b3);
          mma_bf16(o[dp2 * 2], pa[j], b0, b1);
          mma_bf16(o[dp2 * 2 + 1], pa[j], b2, b3);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
      l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
    const int r0 = q0 + g, r1 = r0 + 8;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = nt * 8 + tq * 2;
      if (r0 < t) *reinterpret_cast<uint32_t*>(ob + (long long)r0 * width + c) = pack_bf16(o[nt][0] * inv0, o[nt][1] * inv0);
      if (r1 < t) *reinterpret_cast<uint32_t*>(ob + (long long)r1 * width + c) = pack_bf16(o[nt][2] * inv1, o[nt][3] * inv1);
    }
  }
  trace.end();
}

int vit_attention(vfmreg_ctx* ctx, const __nv_bfloat16* qkv, int b, int t, int heads, int width, __nv_bfloat16* out) {
  VFM_CHECK_ARG(width == heads * ATT_DH, "attention: head dim must be 64 (width %d, heads %d)", width, heads);
  const int tp = (t + 63) / 64 * 64;
  const size_t smem = (size_t)(2 * tp + ATT_WARPS * 16) * ATT_LD * 2;
  VFM_CHECK_ARG(smem <= 200 * 1024, "attention: %d tokens per image do not fit the shared-memory K/V staging", t);
  if (smem > ctx->attention_smem_attr) {   // per device (= per context)
    VFM_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx->attention_smem_attr = smem;
  }
  // query splits: enough CTAs for two per SM, at least one round of 8 query tiles each
  const int n_tiles = (t + 15) / 16;
  int splits = ceil_div(2 * ctx->sm_count, heads * b);
  const int max_splits = (n_tiles + ATT_WARPS - 1) / ATT_WARPS;
  splits = splits < 1 ? 1 : (splits > max_splits ? max_splits : splits);
  VFM_CUDA(launch_pdl(attention_kernel, dim3(heads, b, splits), dim3(ATT_WARPS * 32), smem, ctx->stream, qkv, t, width, tp, out));
  return launch_check(ctx, "attention_kernel");
}

}  // namespace vfm

VFM_TRACE_ATTACH(vfmreg_trace_attach_ops)
