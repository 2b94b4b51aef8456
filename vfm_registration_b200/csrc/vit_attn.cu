// Multi-head self-attention of the DINOv2 forward on tcgen05 (reference: the attention inside the ViT behind
// image_features.py:95-101; xformers memory_efficient_attention there = plain softmax(Q K^T / sqrt(64)) V).
//
// One CTA per (head, image).  K and V of the head (t <= 320 tokens x 64 channels, bf16) are loaded once by TMA straight out
// of the QKV projection's output (row pitch 3 W) into 128B-swizzled shared-memory tiles; the queries are walked in blocks of
// 128 rows:
//   S = Q K^T      tcgen05.mma kind::f16, M128 x N(keys, <= 256 + rest) x K64, both operands K-major, accumulator in TMEM
//                  (columns 0 .. keys)
//   softmax        8 warps, two per TMEM lane quarter: a thread owns one query row (its TMEM lane) and half of the keys:
//                  tcgen05.ld 32x32b.x32 chunks (the next one in flight while one is processed), max pass, maxima of the two
//                  halves exchanged through shared memory, then p = exp2((s - max) / sqrt(64) * log2 e) summed in fp32 and
//                  written as bf16 into a 128B-swizzled K-major tile P in shared memory (generic -> async proxy fence
//                  before the MMA reads it)
//   O = P V        tcgen05.mma M128 x N64 x K(keys): A = P from shared memory, B = the V tile exactly as TMA wrote it
//                  (key rows of 128 B) described as an MN-major operand -- no transpose anywhere; accumulator in TMEM
//                  columns 448 .. 511
//   epilogue       the same warps read O (32 channels each), multiply by 1 / (sum of the two halves) and store 64
//                  contiguous bytes per thread.
// Warp 0 (one elected lane, warp-uniform code) issues the TMA loads and the MMAs between its own softmax work (the stages of a
// query block run one after the other anyway: S and P of two blocks do not fit TMEM / shared memory at 257 keys); the S MMAs
// of block i + 1 are issued as soon as every warp holds S(i) in registers and run under the exponentials of block i; Q blocks
// are double buffered; V has its own mbarrier and lands under the first block's softmax.  A last block of at most AT_FEW real
// rows (token 257 of a 16 x 16 patch grid) is computed on CUDA cores by all 256 threads, without S / P V MMAs.
// What bounds it: the latency of the chain S -> max -> exp -> P -> P V -> O inside a block at two warps per scheduler (ncu:
// XU pipe 20 %, issue slots 26 % busy, a third of the warp samples in mbarrier waits); per block of 128 x 257 about 0.8 us of
// TMEM read + max, 1.4 us of exponentials and P stores, 0.6 us of P V (operand-read bound), 0.5 us of epilogue -- measured
// step by step in profiles/r2_vit_attn_steps.txt.  Masked: keys >= t (scores), queries >= t (never stored).  Token counts
// above 320 use attention_kernel (vit_ops.cu, mma.sync).
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "vit.cuh"

namespace vfm {

constexpr int AT_DH = 64, AT_QB = 128;
constexpr int AT_MAX_CH = 5;                  // 32-key chunks per softmax warp (two warps share a row)
constexpr int AT_MAX_T = AT_MAX_CH * 2 * 32;   // 320 tokens per image
constexpr int AT_FEW = 4;                     // a last query block with at most this many real rows is computed without the tensor core
// O = P V is a chain of keys / 16 MMAs of 128 x 64 x 16 into one accumulator: ~70 cycles each, measured (0.6 us for the 17 of
// a 257-token image) against 32 cycles of tensor work.  Not the dependency between them: dealt round-robin over AT_NACC = 3
// independent accumulators (TMEM columns 320 / 384 / 448, added up by the epilogue) the chain took exactly as long and the
// epilogue 0.1 us longer (profiles/r2_vit_attn_steps.txt) -- it is the operand read: 4 KB of P + 2 KB of V out of shared
// memory per MMA.  AT_NACC stays 1.
constexpr int AT_NACC = 1;
constexpr uint32_t AT_O_COL = 512 - AT_NACC * 64;   // first TMEM column of the O accumulators

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// shared memory: [K tile | V tile | Q block x 2 | P tile | barriers]; tiles are 1024-byte aligned (swizzle atoms)
struct AttnSmem {
  uint32_t k, v, q, p, bars;
  uint32_t total;
};
__host__ __device__ inline AttnSmem attn_smem(int t) {
  const uint32_t rows64 = (uint32_t)(t + 63) / 64 * 64;     // K / V rows loaded (boxes of 64 rows)
  const uint32_t kblocks = (uint32_t)(t + 63) / 64;          // P: 64-key blocks of 128 rows x 128 B
  AttnSmem s;
  s.k = 0;
  s.v = s.k + rows64 * 128;
  s.q = s.v + rows64 * 128;
  s.p = s.q + 2 * AT_QB * 128;
  s.bars = s.p + kblocks * AT_QB * 128;
  s.total = s.bars + 128 + 4 * AT_QB * 4 + AT_FEW * AT_MAX_T * 4 + 64 + 1024;   // barriers, row max / sum exchange, shared rows, slack
  return s;
}

__global__ void __launch_bounds__(256, 1)
    attention_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, int t, int width, __nv_bfloat16* __restrict__ out,
                        const char* pf_ptr, unsigned long long pf_bytes) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const AttnSmem L = attn_smem(t);
  const uint32_t sK = base + L.k, sV = base + L.v, sQ = base + L.q, sP = base + L.p, bars = base + L.bars;
  const uint32_t bar_kv = bars, bar_q0 = bars + 8, bar_s = bars + 24, bar_p = bars + 32, bar_o = bars + 40, bar_sfree = bars + 48, bar_v = bars + 56;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bars + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool leader = threadIdx.x == 0;             // barrier init, trace marks
  // Warp 0 issues every TMA load and MMA of the CTA, as a WHOLE warp with one elected lane: the descriptors are then
  // warp-uniform values in uniform registers.  Issued from a single-thread branch (threadIdx.x == 0) every tcgen05.mma was
  // wrapped by the compiler in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop of ~28 dependent instructions -- 58 ns per MMA,
  // 1 us for the 17 MMAs of one P V product, on the critical path of every query block (profiles/r2_vit_attn_steps.txt).
  const bool issuer = warp == 0;
  const int head = blockIdx.x, img = blockIdx.y;
  const int kp = (t + 15) / 16 * 16;                // keys covered by the MMAs (scores of keys >= t are masked, their P is 0)
  const int n_qb = (t + AT_QB - 1) / AT_QB;
  const int n_acc = kp / 16 < AT_NACC ? kp / 16 : AT_NACC;
  const int row0 = img * t;                         // first token row of this image in the QKV matrix
  const int rows64 = (t + 63) / 64 * 64;

  TraceScope trace(11);
  TraceMarks marks;   // phases of CTA 0, thread 0
  pdl_launch_dependents();
  if (issuer && elect_one()) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qkv) : "memory");
    mbar_init(bar_kv, 1);   // K tile (S needs Q and K only; V has its own barrier and lands under the first block's softmax)
    mbar_init(bar_v, 1);
    mbar_init(bar_q0, 1);
    mbar_init(bar_q0 + 8, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 8);
    mbar_init(bar_o, 1);
    mbar_init(bar_sfree, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  trace.waited();

  if (threadIdx.x == 224)   // a thread of the last warp: off the issuer's path to the K / V / Q loads
    l2_prefetch_slice(pf_ptr, pf_bytes, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y);
  const int n_lo = kp < 256 ? kp : 256, n_hi = kp - n_lo;
  const uint32_t idesc_lo = umma_idesc_f16(AT_QB, n_lo, 1), idesc_hi = umma_idesc_f16(AT_QB, n_hi > 0 ? n_hi : 16, 1);
  const uint32_t idesc_pv = umma_idesc_f16(AT_QB, AT_DH, 1) | (1u << 16);   // B (= V) is MN-major: channels contiguous
  const uint64_t dK = umma_desc_k_sw128(sK), dV = umma_desc_k_sw128(sV), dP = umma_desc_k_sw128(sP);
  auto load_q = [&](int qb) {   // one elected lane: query block qb -> buffer qb & 1
    const uint32_t nb = (uint32_t)(qb & 1);
    mbar_expect_tx(bar_q0 + 8 * nb, AT_QB * 128u);
    tma_load_2d(sQ + nb * (AT_QB * 128), &map_qkv, bar_q0 + 8 * nb, head * AT_DH, row0 + qb * AT_QB);
    tma_load_2d(sQ + nb * (AT_QB * 128) + 64 * 128, &map_qkv, bar_q0 + 8 * nb, head * AT_DH, row0 + qb * AT_QB + 64);
  };
  auto issue_s = [&](int qb) {   // warp 0, converged: S = Q(qb) K^T into TMEM columns 0 .. kp
    const uint32_t qbuf = (uint32_t)(qb & 1);
    mbar_wait(bar_q0 + 8 * qbuf, (uint32_t)((qb >> 1) & 1));
    tc_fence_after();
    const uint64_t dQ = umma_desc_k_sw128(sQ + qbuf * (AT_QB * 128));
    if (elect_one()) {
#pragma unroll
      for (int k = 0; k < AT_DH / 16; ++k) {
        tc_mma_f16(tmem_base, dQ + (uint64_t)(k * 2), dK + (uint64_t)(k * 2), idesc_lo, k != 0 ? 1u : 0u);
        if (n_hi > 0)   // keys 256 ..: rows 256.. of the K tile = +32 KB
          tc_mma_f16(tmem_base + 256, dQ + (uint64_t)(k * 2), dK + (uint64_t)((256 * 128) >> 4) + (uint64_t)(k * 2), idesc_hi,
                     k != 0 ? 1u : 0u);
      }
      tc_commit(bar_s);
    }
    __syncwarp();
  };
  if (issuer) {
    if (elect_one()) {
      mbar_expect_tx(bar_kv, (uint32_t)rows64 * 128u);
      mbar_expect_tx(bar_v, (uint32_t)rows64 * 128u);
      load_q(0);   // Q first: the S MMAs need Q and K
      for (int r = 0; r < rows64; r += 64) tma_load_2d(sK + r * 128, &map_qkv, bar_kv, width + head * AT_DH, row0 + r);
      for (int r = 0; r < rows64; r += 64) tma_load_2d(sV + r * 128, &map_qkv, bar_v, 2 * width + head * AT_DH, row0 + r);
      if (n_qb > 1) load_q(1);
    }
    __syncwarp();
    mbar_wait(bar_kv, 0);
    if (leader) marks.mark(100);
    issue_s(0);
    if (leader) marks.mark(101);
  }

  // ===== softmax + epilogue: all 8 warps.  TMEM lane quarter = warp % 4 (a hardware rule); the two warps of a quarter
  // (warp, warp + 4) own the same 32 query rows and split the keys in 32-wide chunks (the second takes the odd one) and the
  // 64 output channels; row maximum and row sum are exchanged through shared memory. =====
  const int q4 = warp & 3, half = warp >> 2;
  const int r = q4 * 32 + lane;                    // row inside the query block = TMEM lane
  const uint32_t t_lane = tmem_base + ((uint32_t)(q4 * 32) << 16);
  const float sc = 0.125f * 1.4426950408889634f;   // 1 / sqrt(64) * log2(e)
  const uint32_t p_row = sP + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
  float* xmax = reinterpret_cast<float*>(smem + L.bars + 128);   // [2][128]
  float* xsum = xmax + 2 * AT_QB;                                  // [2][128]
  // the second warp of a pair takes the odd chunk: it is the one with the partial last chunk (257 tokens: 4 full chunks against
  // 4 full + 1 key), and warp 0 of the first halves also issues the MMAs
  const int n_chunks = (kp + 31) >> 5, c_split = n_chunks >> 1;
  const int c_begin = half ? c_split : 0, c_end = half ? n_chunks : c_split;
  const int pair_bar = 1 + q4;   // named barrier of the two warps of this quarter (0 is __syncthreads)
  for (int qb = 0; qb < n_qb; ++qb) {
    const int q_row = qb * AT_QB + r;
    const bool warp_live = qb * AT_QB + q4 * 32 < t;   // warp-uniform: some row of this warp is a real query
    const int live_rows = t - qb * AT_QB;   // real query rows of this block (>= 128: all)
    if (live_rows <= AT_FEW) {
      // ---- the last block when it holds a handful of real rows (the 257th token of a 16 x 16 patch grid).  As a 128-row block
      // it cost 2.7 us of the CTA's 15.8 (S and P V MMAs, the exponentials of one thread per row); here no tensor core is
      // involved: all 256 threads share each row -- thread j scores keys j and j + 256 against the K tile in shared memory,
      // the block reduces maximum and sum, and P V is 8 key slices x 32 channel pairs over the V tile.
      float* srow = xsum + 2 * AT_QB;   // [AT_MAX_T] probabilities of the row (bf16-rounded, like the P tile of a full block)
      float* wpart = srow + AT_FEW * AT_MAX_T;   // [8] per-warp partials
      float* obuf = xmax;   // [8][64] partial outputs (the row max / sum exchange area of the full blocks: 512 floats)
      const int tid = threadIdx.x;
      mbar_wait(bar_kv, 0);   // every thread reads the K / V tiles itself here: each orders its reads behind the TMA writes
      mbar_wait(bar_v, 0);
      mbar_wait(bar_q0 + 8 * (uint32_t)(qb & 1), (uint32_t)((qb >> 1) & 1));   // the block's Q rows have landed
      auto lo = [](uint32_t u) { return __uint_as_float(u << 16); };
      auto hi = [](uint32_t u) { return __uint_as_float(u & 0xffff0000u); };
      for (int row = 0; row < live_rows; ++row) {
        // Q row `row` of the block: 128 B, 16-byte chunks XOR-swizzled with the row index (rows < 8: one swizzle atom)
        const uint8_t* qp = smem + L.q + (uint32_t)(qb & 1) * (AT_QB * 128) + row * 128;
        float qf[AT_DH];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint4 u = *reinterpret_cast<const uint4*>(qp + ((ch ^ (row & 7)) << 4));
          qf[8 * ch + 0] = lo(u.x); qf[8 * ch + 1] = hi(u.x); qf[8 * ch + 2] = lo(u.y); qf[8 * ch + 3] = hi(u.y);
          qf[8 * ch + 4] = lo(u.z); qf[8 * ch + 5] = hi(u.z); qf[8 * ch + 6] = lo(u.w); qf[8 * ch + 7] = hi(u.w);
        }
        auto score = [&](int key) {
          const uint8_t* kr = smem + L.k + key * 128;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const uint4 u = *reinterpret_cast<const uint4*>(kr + ((ch ^ (key & 7)) << 4));
            a0 = fmaf(qf[8 * ch + 0], lo(u.x), a0); a1 = fmaf(qf[8 * ch + 1], hi(u.x), a1);
            a2 = fmaf(qf[8 * ch + 2], lo(u.y), a2); a3 = fmaf(qf[8 * ch + 3], hi(u.y), a3);
            a0 = fmaf(qf[8 * ch + 4], lo(u.z), a0); a1 = fmaf(qf[8 * ch + 5], hi(u.z), a1);
            a2 = fmaf(qf[8 * ch + 6], lo(u.w), a2); a3 = fmaf(qf[8 * ch + 7], hi(u.w), a3);
          }
          return (a0 + a1) + (a2 + a3);
        };
        const float s0 = tid < t ? score(tid) : -INFINITY;
        const float s1 = tid + 256 < t ? score(tid + 256) : -INFINITY;
        float mx = fmaxf(s0, s1);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        if (lane == 0) wpart[warp] = mx;
        asm volatile("bar.sync 9, 256;" ::: "memory");
        mx = wpart[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) mx = fmaxf(mx, wpart[w]);
        const float mb = -mx * sc;
        const float p0 = tid < t ? ex2_approx(fmaf(s0, sc, mb)) : 0.f;
        const float p1 = tid + 256 < t ? ex2_approx(fmaf(s1, sc, mb)) : 0.f;
        float sum = p0 + p1;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (tid < t) srow[tid] = __bfloat162float(__float2bfloat16(p0));
        if (tid + 256 < t) srow[tid + 256] = __bfloat162float(__float2bfloat16(p1));
        asm volatile("bar.sync 9, 256;" ::: "memory");   // every thread has read the maxima
        if (lane == 0) wpart[warp] = sum;
        // O[row][2 c2, 2 c2 + 1] over the keys = part (mod 8): V rows are 128 B, chunks XOR-swizzled with key % 8 = part
        const int c2 = tid & 31, part = tid >> 5;
        const uint8_t* vp = smem + L.v + ((((c2 >> 2) ^ part) << 4) + (c2 & 3) * 4);
        float o0 = 0.f, o1 = 0.f;
#pragma unroll 8
        for (int key = part; key < t; key += 8) {   // unrolled: the loads of 8 keys in flight, not one round trip per key
          const uint32_t u = *reinterpret_cast<const uint32_t*>(vp + key * 128);
          const float pk = srow[key];
          o0 = fmaf(pk, lo(u), o0);
          o1 = fmaf(pk, hi(u), o1);
        }
        obuf[part * AT_DH + 2 * c2] = o0;
        obuf[part * AT_DH + 2 * c2 + 1] = o1;
        asm volatile("bar.sync 9, 256;" ::: "memory");
        if (tid < AT_DH) {
          float tot = wpart[0], o = obuf[tid];
#pragma unroll
          for (int w = 1; w < 8; ++w) {   // fixed order: deterministic
            tot += wpart[w];
            o += obuf[w * AT_DH + tid];
          }
          out[(long long)(row0 + qb * AT_QB + row) * width + head * AT_DH + tid] = __float2bfloat16(o / tot);
        }
        asm volatile("bar.sync 9, 256;" ::: "memory");   // srow / wpart / obuf are free for the next row
      }
      break;
    }
    mbar_wait(bar_s, (uint32_t)(qb & 1));
    if (leader) marks.mark(110);
    tc_fence_after();
    // the scores of this thread's keys are read from TMEM ONCE and stay in registers (up to AT_MAX_CH chunks): a second pass
    // over the 128 x keys fp32 tile would add another ~0.5 us of TMEM reads per block
    uint32_t v[AT_MAX_CH][32];
    if (warp_live) {
#pragma unroll
      for (int i = 0; i < AT_MAX_CH; ++i)
        if (c_begin + i < c_end) tc_ld32(t_lane + (c_begin + i) * 32, v[i]);
      tc_wait_ld();
    }
    // S(qb) now lives in registers: its TMEM columns are free for S(qb + 1), which warp 0 issues as soon as all 8 warps
    // are here -- those MMAs then run under the exponentials instead of between the P V product and the next block
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_sfree);
    if (warp_live) {
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < AT_MAX_CH; ++i)
        if (c_begin + i < c_end) {
          const int valid = t - (c_begin + i) * 32;   // keys of the chunk are real iff index < valid
          if (valid >= 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[i][j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < valid) mx = fmaxf(mx, __uint_as_float(v[i][j]));
          }
        }
      xmax[half * AT_QB + r] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      mx = fmaxf(xmax[r], xmax[AT_QB + r]);
      if (leader) marks.mark(111);
      if (issuer && t - (qb + 1) * AT_QB > AT_FEW) {   // the next block exists and is not on the few-rows path
        mbar_wait(bar_sfree, (uint32_t)(qb & 1));
        issue_s(qb + 1);
        if (qb + 2 < n_qb) {   // into the buffer S(qb) read; those MMAs completed before bar_s(qb)
          if (elect_one()) load_q(qb + 2);
          __syncwarp();
        }
        if (leader) marks.mark(106);
      }
      const float mb = -mx * sc;
      // ---- p = exp2((s - max) * sc) -> row sum (fp32) and bf16 P tile.  ~1.4 us per block for a warp's 4 - 5 chunks: neither
      // the MUFU pipe (evaluating every second exponential by a polynomial on the FMA pipe changed nothing) nor the issue
      // slots (20 % busy) bound it, but the latency of two warps per scheduler at 255 registers; issuing the 32 MUFU.EX2 of a
      // chunk back to back in a pass of their own was slower (profiles/r2_vit_attn_steps.txt).
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < AT_MAX_CH; ++i)
        if (c_begin + i < c_end) {
          const int c = c_begin + i;
          const int valid = t - c * 32;
          uint32_t pk[16];
          if (valid >= 32) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float p0 = ex2_approx(fmaf(__uint_as_float(v[i][j]), sc, mb));
              const float p1 = ex2_approx(fmaf(__uint_as_float(v[i][j + 1]), sc, mb));
              s0 += p0;
              s1 += p1;
              pk[j >> 1] = pack_bf16x2(p0, p1);
            }
            sum += s0 + s1;
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float p0 = j < valid ? ex2_approx(fmaf(__uint_as_float(v[i][j]), sc, mb)) : 0.f;
              const float p1 = j + 1 < valid ? ex2_approx(fmaf(__uint_as_float(v[i][j + 1]), sc, mb)) : 0.f;
              sum += p0 + p1;
              pk[j >> 1] = pack_bf16x2(p0, p1);
            }
          }
          // 32 keys = four 16-byte chunks of this row in key block c / 2, XOR-swizzled with the row index (SWIZZLE_128B)
          const uint32_t blk = p_row + (uint32_t)(c >> 1) * (AT_QB * 128);
          const int ch0 = (c & 1) << 2;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            if (c * 32 + 8 * k4 < kp) {
              const uint32_t addr = blk + (uint32_t)(((ch0 + k4) ^ (r & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * k4]), "r"(pk[4 * k4 + 1]),
                           "r"(pk[4 * k4 + 2]), "r"(pk[4 * k4 + 3])
                           : "memory");
            }
          }
        }
      xsum[half * AT_QB + r] = sum;
    }
    if (leader) marks.mark(112);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // P (generic-proxy stores) -> visible to the MMA
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);
    if (issuer) {
      // every warp has written its part of P(qb) (and read O(qb - 1)): O = P V
      mbar_wait(bar_p, (uint32_t)(qb & 1));
      mbar_wait(bar_v, 0);   // (completes once, during the first block)
      if (leader) marks.mark(102);
      tc_fence_after();
      if (elect_one()) {
        for (int kk = 0; kk < kp / 16; ++kk) {
          // A: 16 keys of P = 32 B inside the 128 B swizzle row of key block kk / 4; B: 16 key rows of V = 2 x 1024 B
          const uint64_t da = dP + (uint64_t)(((kk >> 2) * (AT_QB * 128)) >> 4) + (uint64_t)((kk & 3) * 2);
          const uint64_t db = dV + (uint64_t)((kk * 2048) >> 4);
          tc_mma_f16(tmem_base + AT_O_COL + (uint32_t)(kk % AT_NACC) * 64u, da, db, idesc_pv, kk >= AT_NACC ? 1u : 0u);
        }
        tc_commit(bar_o);
      }
      __syncwarp();
      if (leader) marks.mark(104);
#ifdef VFM_ATTN_PROBE   // trace experiment: how long after the last issue do the P V MMAs complete?
      mbar_wait(bar_o, (uint32_t)(qb & 1));
      if (leader) marks.mark(105);
#endif
      if (leader) marks.mark(103);
    }
    mbar_wait(bar_o, (uint32_t)(qb & 1));
    if (leader) marks.mark(113);
    tc_fence_after();
    if (warp_live) {
      uint32_t o[32];
      tc_ld32(t_lane + AT_O_COL + half * 32, o);
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");   // both partial sums of the row are in shared memory
      const float inv_sum = 1.f / (xsum[r] + xsum[AT_QB + r]);
      tc_wait_ld();
#pragma unroll
      for (int a = 1; a < AT_NACC; ++a)
        if (a < n_acc) {   // warp-uniform: accumulators the chain reached (all of them from 48 keys on)
          uint32_t o2[32];
          tc_ld32(t_lane + AT_O_COL + a * 64 + half * 32, o2);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) + __uint_as_float(o2[j]));
        }
      if (q_row < t) {
        __nv_bfloat16* dst = out + (long long)(row0 + q_row) * width + head * AT_DH + half * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[8 * i]) * inv_sum, __uint_as_float(o[8 * i + 1]) * inv_sum);
          w.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv_sum, __uint_as_float(o[8 * i + 3]) * inv_sum);
          w.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv_sum, __uint_as_float(o[8 * i + 5]) * inv_sum);
          w.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv_sum, __uint_as_float(o[8 * i + 7]) * inv_sum);
          *reinterpret_cast<uint4*>(dst + 8 * i) = w;
        }
      }
    }
    if (leader) marks.mark(114);
    // O(qb) has been read by this warp before it takes part in P(qb + 1): PV(qb + 1) is issued after all 8 warps arrived
    tc_fence_before();
  }
  if (leader) marks.flush();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
  trace.end();
}

bool vit_attention_tc_supported(int t) { return t >= 16 && t <= AT_MAX_T; }

int vit_attention_tc(vfmreg_ctx* ctx, const CUtensorMap& map_qkv, int b, int t, int heads, int width, __nv_bfloat16* out,
                     const void* pf_ptr, size_t pf_bytes) {
  VFM_CHECK_ARG(width == heads * AT_DH && vit_attention_tc_supported(t), "attention_tc: unsupported shape (t %d, width %d, heads %d)", t,
                width, heads);
  const size_t smem = attn_smem(t).total;
  if (smem > ctx->attention_tc_smem_attr) {   // per device (= per context)
    VFM_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx->attention_tc_smem_attr = smem;
  }
  VFM_CUDA(launch_pdl(attention_tc_kernel, dim3(heads, b), dim3(256), smem, ctx->stream, map_qkv, t, width, out, (const char*)pf_ptr,
                      (unsigned long long)pf_bytes));
  return launch_check(ctx, "attention_tc_kernel");
}

}  // namespace vfm

VFM_TRACE_ATTACH(vfmreg_trace_attach_attn)
