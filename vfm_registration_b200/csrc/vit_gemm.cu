// bf16 x bf16 -> fp32 GEMM on tcgen05 for the DINOv2 forward (reference: the ViT behind image_features.py:95-101),
// out[token, feature] = X[token, :] . W[feature, :] with the layer's element-wise tail fused into the epilogue so that no
// intermediate is re-read from HBM:
//   EPI_BF16_BIAS       out_bf16 = acc + bias                       (QKV projection)
//   EPI_BF16_BIAS_GELU  out_bf16 = gelu(acc + bias), exact erf      (MLP fc1)
//   EPI_F32_PARTIAL     ws[split][token][feature] = acc, fp32       (attention proj / MLP fc2: the K range may be split over
//                       several clusters; bias, LayerScale, residual add and the reduction over the splits are done by the
//                       LayerNorm kernel that follows, vit_ops.cu, in a fixed order -- deterministic, and the GEMM's
//                       epilogue is a plain coalesced store)
//   EPI_F32_RESID       x[token][feature] += scale * (acc + bias)   (proj / fc2 when the plan does not split K: the residual
//                       add rides in the epilogue, under the tile's MMAs, and the LayerNorm that follows only reads x --
//                       8 B per element less HBM traffic per residual branch than partial + fold; many-image batches)
//   EPI_F32_PATCH       x[token row] = acc + bias + pos_embed       (patch embedding; skips the CLS row of each image)
//
// Layout of the work ("swap-AB"): the WEIGHTS are the M operand of the MMA and the TOKENS the N operand.  A CTA pair
// (thread-block cluster of two, one tcgen05.mma.cta_group::2 issued by the leader drives the tensor cores of both SMs)
// owns 256 output features -- 128 per CTA, accumulator rows in that CTA's TMEM lanes -- times a token tile whose width is a
// run-time value (any multiple of 16 up to 256: the N field of the instruction descriptor).  Why this way round:
//   * the token count of a ViT batch is never a multiple of 128 (B x 257 with the CLS rows: 1542 at BASELINE config 3), the
//     feature counts always are.  Token tiles of a width chosen per launch plus one narrow tail tile (16-wide for the 6 rows
//     past 6 x 256) fill the 74 CTA pairs in one or two even waves instead of 13 row blocks x N tiles with a 6-row block;
//   * an epilogue thread owns one FEATURE (its TMEM lane) and walks the tokens: bias / LayerScale are one register each, and the
//     32 lanes of a warp write 32 consecutive features of one token -- coalesced 64 B (bf16) / 128 B (fp32) rows;
//   * each CTA loads only half of the token tile (the pair's MMA reads both halves), so the shared-memory operand traffic
//     per MMA is 4 KB + N/2 x 32 B per SM, under the 128 B/clk limit down to N = 96.
// Persistent clusters walk the tile list (full token tiles first, the narrow tail tiles last) round-robin.  Warp 0 = TMA
// producer (128B-swizzled 64-wide K chunks, 6-stage mbarrier ring; both CTAs' loads credit the leader's barrier), warp 1 =
// TMEM allocator and, in the leader, the elected-lane MMA issuer (multicast tcgen05.commit frees the stage / publishes the
// accumulator in both CTAs), warps 2-9 = epilogue (tcgen05.ld 32x32b.x32; warps w and w+4 share a TMEM lane quarter and
// take alternate 32-token chunks), accumulators double-buffered in TMEM (2 x 256 columns).
// Every kernel of the forward is launched with programmatic stream serialisation (PDL): barrier init, TMEM allocation,
// tensor-map prefetch AND the weight loads of the first ring of stages of launch i+1 overlap the tail of launch i (the weights
// do not depend on it); griddepcontrol.wait precedes the first access to anything another kernel produces.
// Bound: the L2 -> SM operand traffic (64 B/clk per SM for a 256 x 256 x 64 step against ~42 B/clk chip-wide), then the tensor
// pipe -- ncu: 82 % tensor pipe active at 12.8 TB/s of L2 traffic at 48 images; at BASELINE config 3 (6 images) one or two
// waves per GEMM and fill / drain latency.  DESIGN.md 4.3.
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "vit.cuh"

namespace vfm {

#ifndef VFM_GEMM_CLUSTER
#define VFM_GEMM_CLUSTER 2
#endif
// CTAs per cluster: 2 = one CTA pair; 4 = two pairs on adjacent 256-feature blocks that SHARE the token tile (pair 0 loads it
// and multicasts each half to the matching CTA of pair 1: the operand traffic per pair drops from 64 to 48 KB per 64-wide K step)
constexpr int G_CL = VFM_GEMM_CLUSTER;
static_assert(G_CL == 2 || G_CL == 4, "cluster of one or two CTA pairs");
constexpr int G_FM = 128;            // features per CTA (256 per pair)
constexpr int G_BK = 64, G_STAGES = 6;
#ifndef VFM_GEMM_EPI_WARPS
#define VFM_GEMM_EPI_WARPS 8
#endif
constexpr int G_EPI_WARPS = VFM_GEMM_EPI_WARPS, G_THREADS = 64 + G_EPI_WARPS * 32;   // warps sharing a TMEM lane quarter take chunks round-robin
constexpr uint32_t G_A_BYTES = G_FM * G_BK * 2;        // 16 KB: this CTA's 128 weight rows x 64 k
constexpr uint32_t G_B_BYTES = 128 * G_BK * 2;         // up to 128 token rows (half of a 256-wide tile) x 64 k
constexpr uint32_t G_STAGE_BYTES = G_A_BYTES + G_B_BYTES;
constexpr uint32_t G_SMEM_BARS = G_STAGES * G_STAGE_BYTES;
constexpr uint32_t G_SMEM_TOTAL = G_SMEM_BARS + 256 + 1024;
static_assert(G_SMEM_TOTAL <= 232448, "shared memory budget of one SM");

// erf-GELU, 0.5 x (1 + erf(x / sqrt 2)).  libm's erff costs ~25 instructions per element and made the fc1 epilogue the
// limiter of the whole GEMM (ncu: tensor pipe 37 % active at 48 images against 82 % for the QKV projection).  Here
// 1 + erf(z) = erfc(-z) and erfc(a) = 2^(-a q(a)) for a >= 0 with q a degree-4 polynomial (weighted least squares on
// [0, 4], |erf error| < 7.2e-7 in fp32 arithmetic, tools/fit_erf.py; erfc(4) = 1.5e-8 is below half an ulp of 1):
// one MUFU.EX2 and ~11 FMA-pipe instructions, no branch, and full relative accuracy in the negative tail.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = x * 0.70710678118654752440f;
  const float a = fminf(fabsf(z), 4.0f);
  float q = 0.002966806525364518f;
  q = fmaf(q, a, -0.02966925874352455f);
  q = fmaf(q, a, 0.148755744099617f);
  q = fmaf(q, a, 0.9184716939926147f);
  q = fmaf(q, a, 1.6278934478759766f);
  float e;                                  // erfc(|z|); the exponent is >= -27: no denormal handling needed
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-a * q));
  const float h = 0.5f * x;
  return h * (z >= 0.f ? 2.0f - e : e);     // 1 + erf(z)
}

// tile t of the launch: full token tiles first (token-tile major, feature block minor), then the tail tiles
struct GemmTile {
  int fb, tok0, w, split, kb0, kb1;
};
__device__ __forceinline__ GemmTile gemm_tile(const GemmEpilogue& ep, int u) {
  GemmTile g;
  const int t = u / ep.split;
  g.split = u - t * ep.split;
  const int kbs = ep.k / G_BK;
  g.kb0 = (kbs * g.split) / ep.split;
  g.kb1 = (kbs * (g.split + 1)) / ep.split;
  const int full = ep.fb_count * ep.n_full;
  if (t < full) {
    g.fb = t % ep.fb_count;
    g.tok0 = (t / ep.fb_count) * ep.nt;
    g.w = ep.nt;
  } else {
    g.fb = t - full;
    g.tok0 = ep.n_full * ep.nt;
    g.w = ep.tail_w;
  }
  return g;
}

// One 32-token chunk of one feature (r[j] = accumulator of token tok_c + j): the epilogue's element-wise tail and store.
template <int EPI, bool FULL>
__device__ __forceinline__ void epilogue_chunk(const GemmEpilogue& ep, const uint32_t* r, int tok_c, int f, float bias, float scale,
                                               int split, int valid) {
  if (EPI == EPI_BF16_BIAS || EPI == EPI_BF16_BIAS_GELU) {
    __nv_bfloat16* o = ep.out_bf16 + (long long)tok_c * ep.ldo + f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (FULL || j < valid) {
        float v = __uint_as_float(r[j]) + bias;
        if (EPI == EPI_BF16_BIAS_GELU) v = gelu_erf(v);
        *o = __float2bfloat16(v);
      }
      o += ep.ldo;
    }
  } else if (EPI == EPI_F32_PARTIAL) {
    float* o = ep.x + ((long long)split * ep.m + tok_c) * ep.ldo + f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (FULL || j < valid) *o = __uint_as_float(r[j]);
      o += ep.ldo;
    }
  } else if (EPI == EPI_F32_RESID) {
    // handled by resid_load / resid_store (the loads of x run a chunk ahead of the accumulator reads)
  } else {
    // patch row tok = img * np + p  ->  residual-stream row img * (np + 1) + 1 + p, position row 1 + p.  (img, p) of every
    // token of the chunk follow from the first one's without a carried dependency: 32 independent load / add / store chains
    const int img0 = tok_c / ep.np, p0 = tok_c - img0 * ep.np;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (FULL || j < valid) {
        int p = p0 + j, img = img0;
        while (p >= ep.np) {   // at most a few images per 32-token chunk
          p -= ep.np;
          ++img;
        }
        const long long out_row = (long long)img * (ep.np + 1) + 1 + p;
        ep.x[out_row * ep.ldo + f] = __uint_as_float(r[j]) + bias + __ldg(ep.pos + (long long)(1 + p) * ep.n + f);
      }
    }
  }
}

// EPI_F32_RESID: x[token][feature] += scale * (acc + bias).  The 32 values of x a thread needs for a chunk do not depend on the
// MMAs, so they are requested before the accumulator is waited for and, inside a tile, one chunk ahead: an epilogue warp
// always has 4 KB of reads in flight instead of taking an L2 / HBM round trip per chunk in the shadow of nothing.
__device__ __forceinline__ int resid_valid(const GemmEpilogue& ep, const GemmTile& g, int c) {
  return min(32, min(g.w - c * 32, ep.m - (g.tok0 + c * 32)));   // warp-uniform
}
__device__ __forceinline__ void resid_load(const GemmEpilogue& ep, const GemmTile& g, int c, int f, float* xv) {
  const int valid = resid_valid(ep, g, c);
  const float* o = ep.x + (long long)(g.tok0 + c * 32) * ep.ldo + f;
  if (valid == 32) {
#pragma unroll
    for (int j = 0; j < 32; ++j) xv[j] = o[(long long)j * ep.ldo];
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) xv[j] = j < valid ? o[(long long)j * ep.ldo] : 0.f;
  }
}
// same expression as the fold in the normalisation kernel (add_residual_n with one split): identical bits either way
__device__ __forceinline__ void resid_store(const GemmEpilogue& ep, const GemmTile& g, int c, int f, const uint32_t* r, const float* xv,
                                            float bias, float scale) {
  const int valid = resid_valid(ep, g, c);
  float* o = ep.x + (long long)(g.tok0 + c * 32) * ep.ldo + f;
  if (valid == 32) {
#pragma unroll
    for (int j = 0; j < 32; ++j) o[(long long)j * ep.ldo] = fmaf(scale, __uint_as_float(r[j]) + bias, xv[j]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < valid) o[(long long)j * ep.ldo] = fmaf(scale, __uint_as_float(r[j]) + bias, xv[j]);
  }
}

template <int EPI>
__global__ void __cluster_dims__(G_CL, 1, 1) __launch_bounds__(G_THREADS, 1)
    vit_gemm_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
                    const __grid_constant__ CUtensorMap map_x_tail, const GemmEpilogue ep) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bars = base + G_SMEM_BARS;
  const uint32_t full0 = bars, empty0 = bars + 8 * G_STAGES, tfull0 = bars + 16 * G_STAGES, tempty0 = tfull0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + G_SMEM_BARS + 16 * G_STAGES + 32);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_cta_rank();
  const uint32_t rank = crank & 1u, pair = crank >> 1, lead = crank & ~1u;   // rank inside the CTA pair, pair inside the cluster, the pair's leader
  const int clusters = gridDim.x / G_CL, cid = blockIdx.x / G_CL;
  const int total = ep.fb_count * (ep.n_full + (ep.tail_w > 0 ? 1 : 0)) * ep.split;   // units of (tile, K split)

  TraceScope trace(EPI);
  pdl_launch_dependents();   // the next kernel may start its own prologue; its global accesses wait for this grid to finish
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    if (ep.tail_w > 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x_tail) : "memory");
    for (int s = 0; s < G_STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);    // leader: one arrive.expect_tx per phase, bytes from both CTAs' loads
      mbar_init(empty0 + 8 * s, G_CL / 2);   // every CTA: multicast commit from the MMA warp of each pair's leader
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 1);                  // both CTAs: multicast commit
      mbar_init(tempty0 + 8 * b, 2 * G_EPI_WARPS);   // leader: epilogue warps of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();   // barriers of both CTAs initialised before any remote arrive / multicast commit / 2-SM load
  if (warp == 1) tmem_alloc_2sm(smem_u32(tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    // ===== TMA producer: this CTA's 128 weight rows and its half of the token tile =====
    // The weights do not depend on the previous kernel: their loads for the first ring of stages (and the L2 prefetch of a
    // later GEMM's weights) are issued BEFORE griddepcontrol.wait, while the previous kernel is still running -- after a
    // LayerNorm (no shared memory) this CTA is resident microseconds early.  Only the token loads wait.
    uint32_t stage = 0, phase = 0;
    int pre = 0;
    if (cid < total) {
      const GemmTile g0 = gemm_tile(ep, cid);
      pre = min(G_STAGES, g0.kb1 - g0.kb0);
      if (elect_one()) {
        const int w_row = ((G_CL / 2) * g0.fb + (int)pair) * 2 * G_FM + (int)rank * G_FM;
        const uint32_t b_bytes = (uint32_t)(g0.w / 2) * (G_BK * 2);
        for (int s = 0; s < pre; ++s) {
          const uint32_t full_leader = mapa_cluster(full0 + 8 * s, lead);
          if (rank == 0) mbar_expect_tx(full0 + 8 * s, 2u * (G_A_BYTES + b_bytes));
          tma_load_2d_2sm(base + s * G_STAGE_BYTES, &map_w, full_leader, (g0.kb0 + s) * G_BK, w_row);
        }
        l2_prefetch_slice(ep.pf_ptr, ep.pf_bytes, blockIdx.x, gridDim.x);
      }
      __syncwarp();
    }
    pdl_wait();   // the previous kernel's results (the token matrix) are visible from here on
    trace.waited();
    for (int t = cid; t < total; t += clusters) {
      const GemmTile g = gemm_tile(ep, t);
      const int w_row = ((G_CL / 2) * g.fb + (int)pair) * 2 * G_FM + (int)rank * G_FM;
      const int x_row = g.tok0 + (int)rank * (g.w / 2);
      const bool tail = g.w != ep.nt;   // the tail tile has its own token map: a box of tail_w / 2 rows
      const CUtensorMap* mx = tail ? &map_x_tail : &map_x;
      const uint32_t b_bytes = (uint32_t)(g.w / 2) * (G_BK * 2);
#pragma unroll 1
      for (int kb = g.kb0; kb < g.kb1; ++kb) {
        const bool early = t == cid && kb - g.kb0 < pre;   // barrier armed and weights already on their way
        if (!early) mbar_wait(empty0 + 8 * stage, phase ^ 1);
        if (elect_one()) {
          const uint32_t full_leader = mapa_cluster(full0 + 8 * stage, lead);
          if (!early) {
            if (rank == 0) mbar_expect_tx(full0 + 8 * stage, 2u * (G_A_BYTES + b_bytes));
            tma_load_2d_2sm(base + stage * G_STAGE_BYTES, &map_w, full_leader, kb * G_BK, w_row);
          }
          if (G_CL == 2)
            tma_load_2d_2sm(base + stage * G_STAGE_BYTES + G_A_BYTES, mx, full_leader, kb * G_BK, x_row);
          else if (pair == 0)   // this half of the token tile also goes to the CTA of the same rank in the other pair
            tma_load_2d_2sm_mc(base + stage * G_STAGE_BYTES + G_A_BYTES, mx, full_leader, kb * G_BK, x_row, (uint16_t)(0x5u << rank));
        }
        __syncwarp();
        if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the leader CTA only =====
    pdl_wait();
    if (rank == 0) {   // the leader of each pair
      uint32_t stage = 0, phase = 0;
      int it = 0;
      const uint64_t da0 = umma_desc_k_sw128(base), db0 = umma_desc_k_sw128(base + G_A_BYTES);
      for (int t = cid; t < total; t += clusters, ++it) {
        const GemmTile g = gemm_tile(ep, t);
        const uint32_t idesc = umma_idesc_f16(2 * G_FM, g.w, 1);   // bf16 operands, M = 256 (pair), N = token tile width
        const uint32_t buf = (uint32_t)(it & 1);
        mbar_wait(tempty0 + 8 * buf, (uint32_t)((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 256;
#pragma unroll 1
        for (int kb = g.kb0; kb < g.kb1; ++kb) {
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = da0 + (uint64_t)(stage * (G_STAGE_BYTES >> 4));
            const uint64_t db = db0 + (uint64_t)(stage * (G_STAGE_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < G_BK / 16; ++k)
              tc_mma_f16_2sm(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb != g.kb0 || k != 0) ? 1u : 0u);
            tc_commit_2sm(empty0 + 8 * stage, (uint16_t)((1u << G_CL) - 1));   // stage free, in every CTA of the cluster, once these MMAs have read it
            if (kb == g.kb1 - 1) tc_commit_2sm(tfull0 + 8 * buf, (uint16_t)(3u << (2 * pair)));   // accumulator halves complete
          }
          __syncwarp();
          if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===== epilogue: this CTA's 128 features x the tile's tokens; thread = feature (TMEM lane), chunks of 32 tokens =====
    pdl_wait();   // the epilogue writes what the previous kernels may still be reading
    const int q = warp & 3, half = (warp - 2) >> 2;
    int it = 0;
    const uint32_t tempty_leader0 = mapa_cluster(tempty0, lead);
    for (int t = cid; t < total; t += clusters, ++it) {
      const GemmTile g = gemm_tile(ep, t);
      const int f = ((G_CL / 2) * g.fb + (int)pair) * 2 * G_FM + (int)rank * G_FM + q * 32 + lane;
      const bool f_live = f < ep.n;
      float bias = 0.f, scale = 1.f;
      if (EPI != EPI_F32_PARTIAL && f_live) bias = __ldg(ep.bias + f);
      if (EPI == EPI_F32_RESID && f_live) scale = __ldg(ep.scale + f);
      const uint32_t buf = (uint32_t)(it & 1);
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256;
      const int n_chunks = (g.w + 31) >> 5;
      if (EPI == EPI_F32_RESID) {
        constexpr int CSTEP = G_EPI_WARPS / 4;
        float xa[32], xb[32];
        // (an L2 prefetch of the tile's remaining x lines at this point changed nothing for proj and cost fc2 5 %: not done)
        if (f_live && half < n_chunks) resid_load(ep, g, half, f, xa);
        mbar_wait(tfull0 + 8 * buf, (uint32_t)((it >> 1) & 1));
        tc_fence_after();
#pragma unroll 1
        for (int c = half; c < n_chunks; c += CSTEP) {
          const bool more = f_live && c + CSTEP < n_chunks;
          if (more) resid_load(ep, g, c + CSTEP, f, xb);
          uint32_t r[32];
          tc_ld32(t_addr + c * 32, r);
          tc_wait_ld();
          if (f_live) resid_store(ep, g, c, f, r, xa, bias, scale);
          if (more) {
#pragma unroll
            for (int j = 0; j < 32; ++j) xa[j] = xb[j];
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader0 + 8 * buf);
        continue;
      }
      mbar_wait(tfull0 + 8 * buf, (uint32_t)((it >> 1) & 1));
      tc_fence_after();
#pragma unroll 1
      for (int c = half; c < n_chunks; c += G_EPI_WARPS / 4) {
        uint32_t r[32];
        tc_ld32(t_addr + c * 32, r);
        tc_wait_ld();
        const int tok_c = g.tok0 + c * 32;
        const int valid = min(32, min(g.w - c * 32, ep.m - tok_c));   // warp-uniform
        if (!f_live || valid <= 0) continue;
        // full chunks run straight-line code (32 independent elements in flight); only the last chunk of a tail tile /
        // of the token range takes the predicated path
        if (valid == 32)
          epilogue_chunk<EPI, true>(ep, r, tok_c, f, bias, scale, g.split, 32);
        else
          epilogue_chunk<EPI, false>(ep, r, tok_c, f, bias, scale, g.split, valid);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader0 + 8 * buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs are done with the pair's TMEM and with each other's barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512u);
  }
  trace.end();
}

// ---- host side --------------------------------------------------------------------------------------------------------
// Cost model of one launch (SM cycles): the units (tile x K split) are dealt round-robin to the clusters; a unit costs its
// MMAs -- (k-blocks) x 4 x w/2 cycles for a w-token tile -- or the time its operand bytes need to reach the pair's shared
// memory, whichever is larger, plus a fixed fill / drain; the launch can also not be faster than its total operand bytes
// through the L2 (the chip-wide L2 -> SM rate, ~6300 B/clk, is what bounds 256 x 256 x 64 bf16 tiles: 64 KB per 512 MMA
// cycles per pair = 64 B/clk per SM against ~42 B/clk per SM when all 148 SMs stream).
static long long gemm_cost(int m, int n, int k, int nt, int split, int clusters, GemmPlan* plan) {
  const int fb_count = ceil_div(n, G_CL * G_FM);   // feature blocks of the cluster: 256 per CTA pair
  const int n_full = m / nt;
  const int rest = m - n_full * nt;
  const int tail = (rest + 15) / 16 * 16;
  const int kbs = k / G_BK;
  auto unit_cost = [&](int w, int s, long long* bytes) {
    const int kb = (kbs * (s + 1)) / split - (kbs * s) / split;
    const long long b = (long long)kb * (G_CL * G_A_BYTES + 2 * (long long)(w / 2) * (G_BK * 2));   // the token tile is loaded once per cluster
    *bytes = b;
    const long long mma = (long long)kb * 4 * (w / 2), ingest = b / (50 * G_CL);
    return (mma > ingest ? mma : ingest) + 600 + (long long)(w / 64 + 1) * 250;
  };
  const int tiles_full = fb_count * n_full, tiles = tiles_full + (tail ? fb_count : 0);
  const int units = tiles * split;
  long long worst = 0, total_bytes = 0;
  for (int c = 0; c < clusters && c < units; ++c) {
    long long sum = 0;
    for (int u = c; u < units; u += clusters) {
      long long b;
      sum += unit_cost(u / split < tiles_full ? nt : tail, u % split, &b);
      total_bytes += b;
    }
    if (sum > worst) worst = sum;
  }
  long long l2 = total_bytes / 7500;   // calibrated on the plan sweep (tools/vit_plan_sweep.sh, profiles/r2_vit_plan_sweep_b6.txt)
  if (split > 1) l2 += (long long)split * m * n * 8 / 7500;   // the partial sums are written here and read back by the next kernel
  if (plan) {
    plan->nt = nt;
    plan->split = split;
    plan->tail_w = tail;
    plan->n_full = n_full;
    plan->fb_count = fb_count;
    plan->units = units;
  }
  return worst > l2 ? worst : l2;
}

static int gemm_clusters(vfmreg_ctx* ctx);

// VFMREG_VIT_PLAN="qkv:256:1,proj:256:3,fc1:192:1,fc2:256:3,pe:128:1" overrides (tile width : K splits) per GEMM of the forward --
// a tuning aid for tools/bench_kernels.py; unset = the cost model above
static bool plan_override(const char* which, int* nt, int* split) {
  const char* e = getenv("VFMREG_VIT_PLAN");
  if (!e || !which) return false;
  const char* p = strstr(e, which);
  if (!p || p[strlen(which)] != ':') return false;
  return sscanf(p + strlen(which) + 1, "%d:%d", nt, split) == 2;
}

GemmPlan vit_gemm_plan(vfmreg_ctx* ctx, int m, int n, int k, int max_split, const char* which) {
  const int clusters = gemm_clusters(ctx);
  GemmPlan best{};
  long long best_cost = -1;
  int fnt = 0, fsplit = 0;
  if (plan_override(which, &fnt, &fsplit) && fnt >= 32 && fnt <= 256 && fnt % 32 == 0 && fsplit >= 1 && fsplit <= max_split) {
    gemm_cost(m, n, k, fnt, fsplit, clusters, &best);
    return best;
  }
  for (int split = 1; split <= max_split && split <= k / G_BK; ++split)
    for (int nt = 256; nt >= 64; nt -= 32) {   // multiples of 32: the epilogue's chunks never straddle the end of a full tile
      GemmPlan p{};
      const long long c = gemm_cost(m, n, k, nt, split, clusters, &p);
      if (best_cost < 0 || c < best_cost) {
        best = p;
        best_cost = c;
      }
    }
  return best;
}

// clusters of G_CL CTAs (one per SM, ~200 KB of shared memory each) that can be resident at once: GPC boundaries make this
// smaller than sm_count / G_CL (a cluster cannot straddle two GPCs); asked once per context
static int gemm_clusters(vfmreg_ctx* ctx) {
  if (ctx->gemm_clusters > 0) return ctx->gemm_clusters;
  int n = ctx->sm_count / G_CL;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(G_CL * n);
  cfg.blockDim = dim3(G_THREADS);
  cfg.dynamicSmemBytes = G_SMEM_TOTAL;
  if (cudaFuncSetAttribute(vit_gemm_kernel<EPI_BF16_BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM_TOTAL) == cudaSuccess) {
    int q = 0;
    if (cudaOccupancyMaxActiveClusters(&q, vit_gemm_kernel<EPI_BF16_BIAS>, &cfg) == cudaSuccess && q > 0 && q < n) n = q;
  }
  cudaGetLastError();
  if (getenv("VFMREG_VIT_VERBOSE")) fprintf(stderr, "vit_gemm: %d clusters of %d CTAs resident\n", n, G_CL);
  ctx->gemm_clusters = n;
  return n;
}

template <int EPI>
static int launch_gemm(vfmreg_ctx* ctx, const CUtensorMap& w, const CUtensorMap& x, const CUtensorMap& x_tail, const GemmEpilogue& ep,
                       int grid) {
  const uint64_t bit = 1ull << EPI;   // per device (= per context), not per process
  if (!(ctx->gemm_attr_mask & bit)) {
    VFM_CUDA(cudaFuncSetAttribute(vit_gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM_TOTAL));
    ctx->gemm_attr_mask |= bit;
  }
  VFM_CUDA(launch_pdl(vit_gemm_kernel<EPI>, dim3(grid), dim3(G_THREADS), G_SMEM_TOTAL, ctx->stream, w, x, x_tail, ep));
  return launch_check(ctx, "vit_gemm_kernel");
}

int vit_weight_map(CUtensorMap* map, const void* ptr, int n, int k) { return make_tmap_16bit(map, ptr, n, k, k, G_FM, true); }

int vit_token_maps(TokenMaps* t, const void* ptr, int rows, int k, const GemmPlan& plan) {
  VFM_TRY(make_tmap_16bit(&t->full, ptr, rows, k, k, plan.nt / 2, true));
  if (plan.tail_w > 0 && plan.tail_w != plan.nt) return make_tmap_16bit(&t->tail, ptr, rows, k, k, plan.tail_w / 2, true);
  t->tail = t->full;
  return VFMREG_OK;
}

int vit_gemm(vfmreg_ctx* ctx, int epi, const CUtensorMap& w, const TokenMaps& x, const GemmPlan& plan, GemmEpilogue ep) {
  VFM_CHECK_ARG(ep.k % G_BK == 0 && ep.m > 0 && ep.n % 128 == 0 && plan.nt >= 32 && plan.nt <= 256 && plan.nt % 32 == 0 &&
                    plan.split >= 1 && (plan.split == 1 || epi == EPI_F32_PARTIAL),
                "vit_gemm: unsupported shape m=%d n=%d k=%d nt=%d split=%d", ep.m, ep.n, ep.k, plan.nt, plan.split);
  const int clusters = gemm_clusters(ctx);
  ep.nt = plan.nt;
  ep.split = plan.split;
  ep.fb_count = plan.fb_count;
  ep.n_full = plan.n_full;
  ep.tail_w = plan.tail_w;
  const int grid = G_CL * (plan.units < clusters ? plan.units : clusters);
  switch (epi) {
    case EPI_BF16_BIAS: return launch_gemm<EPI_BF16_BIAS>(ctx, w, x.full, x.tail, ep, grid);
    case EPI_BF16_BIAS_GELU: return launch_gemm<EPI_BF16_BIAS_GELU>(ctx, w, x.full, x.tail, ep, grid);
    case EPI_F32_PARTIAL: return launch_gemm<EPI_F32_PARTIAL>(ctx, w, x.full, x.tail, ep, grid);
    case EPI_F32_PATCH: return launch_gemm<EPI_F32_PATCH>(ctx, w, x.full, x.tail, ep, grid);
    case EPI_F32_RESID: return launch_gemm<EPI_F32_RESID>(ctx, w, x.full, x.tail, ep, grid);
  }
  set_error("vit_gemm: bad epilogue %d", epi);
  return VFMREG_ERR_ARG;
}

}  // namespace vfm

VFM_TRACE_ATTACH(vfmreg_trace_attach_gemm)
