// Row L2 renormalisation (faiss fvec_renorm_L2 at reference VoxelHashMap.cpp:474,480) into a zero-padded
// row stride `dp`.  One warp per row, 128-bit loads; HBM-bound: reads 4*d, writes 4*dp bytes per row.
// Canonical order: lane l accumulates the float4 groups 4l.., 128+4l.., ... with fmaf element by
// element, xor-butterfly 16..1, inv = 1/sqrt(s) (IEEE), y = x * inv; rows with s == 0 are copied unchanged.
#include <cuda_fp16.h>

#include "common.cuh"

namespace vfm {

__global__ void __launch_bounds__(256) normalize_rows_kernel(const float* __restrict__ x, int64_t n, int d, int dp,
                                                           int normalize, float* __restrict__ y, __half* __restrict__ yh,
                                                           uint8_t* __restrict__ nz) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* xr = x + row * (int64_t)d;
  float* yr = y + row * (int64_t)dp;
  const bool vec = ((d & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  float inv = 1.0f;
  if (normalize) {
    float acc = 0.0f;
    if (vec) {
      for (int base = 4 * lane; base < d; base += 128) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(xr + base));
        acc = fmaf(v.x, v.x, acc);
        acc = fmaf(v.y, v.y, acc);
        acc = fmaf(v.z, v.z, acc);
        acc = fmaf(v.w, v.w, acc);
      }
    } else {
      for (int base = 4 * lane; base < d; base += 128)
        for (int c = 0; c < 4 && base + c < d; ++c) {
          const float v = __ldg(xr + base + c);
          acc = fmaf(v, v, acc);
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));
    if (acc > 0.0f) inv = __fdiv_rn(1.0f, __fsqrt_rn(acc));
    if (nz && lane == 0) nz[row] = acc > 0.0f ? 1 : 0;
  } else if (nz && lane == 0) {
    nz[row] = 1;
  }
  // dp is a multiple of 4 and y is arena-aligned: vector stores, zero fill of the padding
  for (int base = 4 * lane; base < dp; base += 128) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec && base + 3 < d) {
      v = __ldg(reinterpret_cast<const float4*>(xr + base));
    } else {
      if (base + 0 < d) v.x = __ldg(xr + base + 0);
      if (base + 1 < d) v.y = __ldg(xr + base + 1);
      if (base + 2 < d) v.z = __ldg(xr + base + 2);
      if (base + 3 < d) v.w = __ldg(xr + base + 3);
    }
    v.x = __fmul_rn(v.x, inv);
    v.y = __fmul_rn(v.y, inv);
    v.z = __fmul_rn(v.z, inv);
    v.w = __fmul_rn(v.w, inv);
    *reinterpret_cast<float4*>(yr + base) = v;
    if (yh) {  // fp16 copy (round-to-nearest) for the tcgen05 candidate search
      __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&h0);
      pk.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(yh + row * (int64_t)dp + base) = pk;
    }
  }
}

int normalize_rows(vfmreg_ctx* ctx, const float* x, int64_t n, int d, int dp, int normalize, float* y, void* yh,
                   uint8_t* nz) {
  if (n <= 0) return VFMREG_OK;
  const int warps = 8;
  normalize_rows_kernel<<<ceil_div(n, warps), warps * 32, 0, ctx->stream>>>(x, n, d, dp, normalize, y,
                                                                             static_cast<__half*>(yh), nz);
  return launch_check(ctx, "normalize_rows");
}

}  // namespace vfm
