// Descriptor top-2 search on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), with results bit-identical to the
// exact fp32 kernel (match_simt.cu) and to oracle/c.
//
// Reference: faiss IndexFlatIP::search(k=1) at VoxelHashMap.cpp:486-495 (+ runner-up for the ratio test).
//
// Two phases
//  1. candidate search -- fp16 x fp16 -> fp32 GEMM of the renormalised descriptors, S~ = A~ B~^T, never written to memory.
//     Three kernels share the epilogue:
//       match_tc3_kernel (default)  CTA pairs, ONE tcgen05.mma.cta_group::2 (M256 N256 K16) per step drives both SMs; each
//                                   CTA holds its 128 query rows (resident for D <= 384) and half of the 256-column tile
//       match_tc2_kernel            CTA pairs, two independent M128 MMAs fed by TMA multicast (A/B runs: VFMREG_MATCH_V2=1)
//       match_tc_kernel             single CTAs, both operands streamed (fewer than two row blocks; VFMREG_MATCH_V1=1)
//     Persistent CTAs walk contiguous spans of (row block, 256-column tile) units in row-major order.  Warp roles: warp 0 =
//     TMA producer (cp.async.bulk.tensor, 128B swizzle, mbarrier ring of 64-wide K chunks), warp 1 = TMEM allocator + MMA
//     issuer (one elected lane, accumulators double-buffered in TMEM: 2 x 256 columns), warps 2-5 = epilogue (tcgen05.ld
//     32x32b: one query row per thread).  Each epilogue thread keeps the running approximate best (and runner-up when the
//     caller needs it) of its row and records every column whose approximate score is within `margin` of it in a small
//     shared-memory list (compacted when full); the lists are flushed per (row block, span) "slot".  The row count may
//     live on the device (TcParams.n_dev: the pruned reverse search of register()'s mutual check), and rows may start
//     from a known lower bound of their best score (TcParams.seed).
//  2. re-rank -- rerank_select: drop the candidates below (approx best / runner-up - margin) and list the survivors;
//     rerank_dot: recompute them in the canonical fp32 order (fmaf chain over k ascending, from the fp32 rows; 16 lanes
//     per candidate hand the accumulator on by shuffle); top-1 mode folds the per-row pick into a packed atomic max
//     (rerank_finish), top-2 mode picks per row (rerank_pick), lowest index on ties.  Rows whose list overflowed
//     (pathological ties) are redone by exact_rows_kernel, an exact scan of all columns.
//
// Why the result is exact: |S~ - S| <= eps with eps bounded below; the exact best and runner-up of a span both have
// S~ >= (final approx runner-up of that span) - 2 eps, the recording threshold only ever rises, so both are always in
// the list; margin = 2 eps + slack.  eps for unit-norm rows: fp16 input rounding 2^-10 (1 + 2^-12) ||a|| ||b||
// + fp16 subnormal inputs (< 5e-5) + tensor-core fp32 accumulation (< D 2^-23) + canonical fp32 chain (< D 2^-24)
// < 1.2e-3 for D <= 1024; MARGIN = 3e-3.  Only valid for renormalised inputs, which is what the caller guarantees.
//
// Bound: tensor pipe (2 N M D flop per launch); measured 0.80-0.83 of the cuBLAS bf16 rate on 10k x 50k x 384.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace vfm {

#ifndef VFM_TBN
#define VFM_TBN 256
#endif
#ifndef VFM_EPI_WARPS
#define VFM_EPI_WARPS 4
#endif
constexpr int TBM = 128, TBN = VFM_TBN, TBK = 64, UMMA_K = 16;
constexpr int STAGES = (TBN == 256) ? 4 : 6;
constexpr int NBUF = 512 / TBN;          // accumulator buffers in TMEM
constexpr int CAP = 16;                 // candidate list entries per (row, slot)
constexpr float MARGIN = 3e-3f;
// Epilogue warps: 4 (one per TMEM lane quarter) or 8 (warps w and w+4 share a lane quarter -- the hardware ties lanes
// to warp_id % 4 -- and split the tile's columns; each (row, column group) then keeps its own candidate list, i.e.
// HALVES device slots per (row, CTA span)).  Measured on B200: 4 and 8 warps perform alike (the epilogue cost is the
// candidate bookkeeping, not issue bandwidth), so the default is 4.
constexpr int EPI_WARPS = VFM_EPI_WARPS, EPI_THREADS = EPI_WARPS * 32;
constexpr int HALVES = EPI_WARPS / 4;     // column groups per tile (warps sharing a TMEM lane quarter split the columns)
constexpr int COLS_PER_WARP = TBN / HALVES;
constexpr int TC_THREADS = 64 + EPI_THREADS;
constexpr uint32_t A_STAGE_BYTES = TBM * TBK * 2, B_STAGE_BYTES = TBN * TBK * 2;
constexpr uint32_t SMEM_RING_V = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
constexpr uint32_t SMEM_RING_I = SMEM_RING_V + EPI_THREADS * CAP * 4;
constexpr uint32_t SMEM_BARS = SMEM_RING_I + EPI_THREADS * CAP * 4;
constexpr uint32_t SMEM_TOTAL = SMEM_BARS + 256 + 1024;  // + slack for 1024-byte alignment

constexpr uint32_t IDESC = umma_idesc_f16(TBM, TBN, 0);  // fp16 operands

struct TcParams {
  int n, m;              // queries (rows of a), database size (rows of b)
  int kb;                // dp / 64
  int col_tiles;         // ceil(m / 256)
  long long total_tiles; // row_blocks * col_tiles
  int slots;             // candidate slots per row
  const uint8_t* nz;     // per query row: 0 = all-zero descriptor
  float* cand_v;         // [n][slots][CAP]
  int* cand_i;
  int* cand_n;           // [n][slots]: count | overflow << 30
  float2* slot_top2;     // [n][slots]: approximate (best, runner-up) of the slot's column span
  long long* dbg;        // optional per-CTA cycle counters (tuning aid): [cta][8]
  int experiment;        // tuning aid: 1 = producer stops issuing TMA after the first ring fill (timing only, garbage results)
  int top1;              // 1: only the best match is needed (runner-up value not requested): threshold = best - margin
  const int* n_dev;      // optional: the number of query rows actually present (<= n), produced earlier on the stream
  const float* seed;     // optional (top1 only): per query row a known lower bound of its exact best score (the pruned
                         // reverse search knows <b_j, a_i> = sim01[i]); recording starts at seed - margin instead of -inf
};

// number of query rows of this launch: the host-side bound, or the device-side count when the caller's row list was
// compacted on the GPU (the pruned reverse search of the mutual check)
__device__ __forceinline__ int tc_rows(const TcParams& P) { return P.n_dev ? min(__ldg(P.n_dev), P.n) : P.n; }

__device__ __forceinline__ void flush_slot(const TcParams& P, int n_rows, long long total_tiles, int rb, int tid, int half,
                                           float* ring_v, int* ring_i, int cnt, bool ovf, float best, float second) {
  const int etid = half * 128 + tid;
  const int row = rb * TBM + tid;
  if (row >= n_rows) return;
  // first CTA whose span [total*c/grid, total*(c+1)/grid) contains this row block's first tile
  const long long t0 = (long long)rb * P.col_tiles;
  const long long g = min((long long)gridDim.x, total_tiles);   // CTAs that own a (non-empty) span
  long long c0 = (t0 * g) / total_tiles;
  while ((total_tiles * (c0 + 1)) / g <= t0) ++c0;
  while (c0 > 0 && (total_tiles * c0) / g > t0) --c0;
  const int slot = (int)(blockIdx.x - c0) * HALVES + half;
  const long long o = ((long long)row * P.slots + slot);
  P.cand_n[o] = cnt | (ovf ? (1 << 30) : 0);
  P.slot_top2[o] = make_float2(best, second);
  for (int e = 0; e < cnt; ++e) {
    P.cand_v[o * CAP + e] = ring_v[e * EPI_THREADS + etid];
    P.cand_i[o * CAP + e] = ring_i[e * EPI_THREADS + etid];
  }
}

#ifdef VFM_SCAN_STATS
// tuning aid (-DVFM_SCAN_STATS): warp-level counts of [0] chunks scanned, [1] chunks where some lane cleared its threshold,
// [2] lane-chunks with several qualifying columns, [3] lane-chunks with at least one
__device__ unsigned long long g_scan_stats[4];
#define SCAN_STAT(i, pred)                                                                 \
  do {                                                                                     \
    const unsigned b__ = __ballot_sync(0xffffffffu, (pred));                               \
    if ((threadIdx.x & 31) == 0 && b__) atomicAdd(&g_scan_stats[i], (i) == 3 ? (unsigned long long)__popc(b__) : 1ull); \
  } while (0)
// lane-level variant, safe inside divergent code
#define SCAN_STAT_LANE(i, pred) do { if (pred) atomicAdd(&g_scan_stats[i], 1ull); } while (0)
#else
#define SCAN_STAT(i, pred) do { } while (0)
#define SCAN_STAT_LANE(i, pred) do { } while (0)
#endif

// Record one candidate (approximate score v of column `col`) in the thread's list and raise the recording threshold.
template <bool TOP1>
__device__ __forceinline__ void push_candidate(float v, int col, int tid, float* ring_v, int* ring_i, float& best, float& second,
                                               float& thr, int& cnt, bool& ovf) {
  if (cnt == CAP) {  // compact: keep what is still above the (risen) threshold
    int w = 0;
#pragma unroll 1
    for (int e = 0; e < CAP; ++e) {
      const float ev = ring_v[e * EPI_THREADS + tid];
      const int ei = ring_i[e * EPI_THREADS + tid];
      if (ev > thr) {
        ring_v[w * EPI_THREADS + tid] = ev;
        ring_i[w * EPI_THREADS + tid] = ei;
        ++w;
      }
    }
    cnt = w;
  }
  if (cnt < CAP) {
    ring_v[cnt * EPI_THREADS + tid] = v;
    ring_i[cnt * EPI_THREADS + tid] = col;
    ++cnt;
  } else {
    ovf = true;
  }
  if (v > best) {
    second = best;
    best = v;
  } else if (v > second) {
    second = v;
  }
  thr = (TOP1 ? best : second) - MARGIN;
}

// r[i] for a lane-dependent i without local memory: a 5-level select tree (31 SEL)
__device__ __forceinline__ float pick32(const uint32_t* r, int i) {
  uint32_t a[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = (i & 16) ? r[16 + k] : r[k];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = (i & 8) ? a[8 + k] : a[k];
#pragma unroll
  for (int k = 0; k < 4; ++k) a[k] = (i & 4) ? a[4 + k] : a[k];
#pragma unroll
  for (int k = 0; k < 2; ++k) a[k] = (i & 2) ? a[2 + k] : a[k];
  return __uint_as_float((i & 1) ? a[1] : a[0]);
}

// One 32-column chunk of a query row (r[i] = approximate score of column col_base + c0 + i), registers only.
// Fast path: four group maxima (group q = columns i % 4 == q), their maximum and one compare against the recording
// threshold.  When the maximum clears the threshold, only the groups whose maximum clears it are searched for the
// qualifying columns (a bit mask); the usual case -- exactly one -- is a single push of (mx, position).  Several
// qualifying columns are pushed in ascending column order through a select tree.  `tid` is the thread's index among
// the epilogue threads.
template <bool TOP1>
__device__ __forceinline__ void scan_chunk(const uint32_t* r, int c0, int valid, int col_base, int tid, float* ring_v, int* ring_i,
                                           float& best, float& second, float& thr, int& cnt, bool& ovf) {
  if (c0 >= valid) return;  // warp-uniform
  const bool partial = c0 + 32 > valid;  // warp-uniform: columns >= valid hold zeros (TMA out-of-bounds fill)
  float g[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float m = fmaxf(__uint_as_float(r[q]), __uint_as_float(r[q + 4]));
#pragma unroll
    for (int j = 2; j < 8; ++j) m = fmaxf(m, __uint_as_float(r[4 * j + q]));
    g[q] = m;
  }
  const float mx = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
  SCAN_STAT(0, true);
  if (thr == -INFINITY && !partial) {
    // Nothing recorded yet in this span: seed the threshold from this chunk.  TOP1: its maximum.  Otherwise a lower bound
    // of its second largest value: the smaller of two disjoint group maxima.
    const float seed = TOP1 ? mx : fminf(fmaxf(g[0], g[2]), fmaxf(g[1], g[3]));
    thr = seed - MARGIN;   // strictly below `seed`, so the values that define it are still recorded below
  }
  SCAN_STAT(1, mx > thr);
#if defined(VFM_SCAN_EXP) && VFM_SCAN_EXP == 1
  if (mx > thr) { best = mx; thr = mx - MARGIN; }   // timing experiment: branch + threshold update only
  return;
#endif
  if (mx > thr) {
    uint32_t mask = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (g[q] > thr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) mask |= (__uint_as_float(r[4 * j + q]) > thr) ? (1u << (4 * j + q)) : 0u;
      }
    }
    if (partial) mask &= (1u << (valid - c0)) - 1u;
#if defined(VFM_SCAN_EXP) && VFM_SCAN_EXP == 2
    best = mx; thr = mx - MARGIN; cnt = (cnt + __popc(mask)) & 7;   // timing experiment: mask, no recording
    return;
#endif
    SCAN_STAT_LANE(2, (mask & (mask - 1)) != 0);
    SCAN_STAT_LANE(3, mask != 0);
    if (!partial && (mask & (mask - 1)) == 0) {
      // exactly one qualifying column: it is the chunk maximum (mask != 0 because mx > thr and every column is valid)
      push_candidate<TOP1>(mx, col_base + c0 + __ffs(mask) - 1, tid, ring_v, ring_i, best, second, thr, cnt, ovf);
    } else {
#pragma unroll 1
      while (mask) {
        const int i = __ffs(mask) - 1;
        mask &= mask - 1;
        const float v = pick32(r, i);
        if (v > thr) push_candidate<TOP1>(v, col_base + c0 + i, tid, ring_v, ring_i, best, second, thr, cnt, ovf);
      }
    }
  }
}

// The epilogue of one warp's share of an accumulator tile (COLS_PER_WARP columns starting at TMEM address t_addr /
// database column col_base): 32-column chunks, the next one in flight while the current one is scanned.
template <bool TOP1>
__device__ __forceinline__ void epilogue_half(uint32_t t_addr, int col_base, int m, int etid, float* ring_v, int* ring_i,
                                              float& best, float& second, float& thr, int& cnt, bool& ovf) {
  const int valid = min(COLS_PER_WARP, m - col_base);   // may be <= 0 for the padded part of the last tile
  uint32_t ra[32], rbuf[32];
  tc_ld32(t_addr, ra);
#pragma unroll 1
  for (int c = 0; c < COLS_PER_WARP / 32; c += 2) {
    tc_wait_ld();
    tc_ld32(t_addr + (c + 1) * 32, rbuf);  // in flight while chunk c is scanned
    scan_chunk<TOP1>(ra, c * 32, valid, col_base, etid, ring_v, ring_i, best, second, thr, cnt, ovf);
    tc_wait_ld();
    if (c + 2 < COLS_PER_WARP / 32) tc_ld32(t_addr + (c + 2) * 32, ra);
    scan_chunk<TOP1>(rbuf, (c + 1) * 32, valid, col_base, etid, ring_v, ring_i, best, second, thr, cnt, ovf);
  }
}

// timing experiment (garbage results): the fast path of the scan only (double-buffered loads + max tree), no recording
__device__ __forceinline__ void epilogue_max_only(uint32_t t_addr, float& best) {
  uint32_t ra[32], rbuf[32];
  tc_ld32(t_addr, ra);
#pragma unroll 1
  for (int c = 0; c < COLS_PER_WARP / 32; c += 2) {
    tc_wait_ld();
    tc_ld32(t_addr + (c + 1) * 32, rbuf);
    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      m0 = fmaxf(m0, __uint_as_float(ra[i]));
      m1 = fmaxf(m1, __uint_as_float(ra[i + 1]));
      m2 = fmaxf(m2, __uint_as_float(ra[i + 2]));
      m3 = fmaxf(m3, __uint_as_float(ra[i + 3]));
    }
    best = fmaxf(best, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
    tc_wait_ld();
    if (c + 2 < COLS_PER_WARP / 32) tc_ld32(t_addr + (c + 2) * 32, ra);
    m0 = m1 = m2 = m3 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      m0 = fmaxf(m0, __uint_as_float(rbuf[i]));
      m1 = fmaxf(m1, __uint_as_float(rbuf[i + 1]));
      m2 = fmaxf(m2, __uint_as_float(rbuf[i + 2]));
      m3 = fmaxf(m3, __uint_as_float(rbuf[i + 3]));
    }
    best = fmaxf(best, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
  }
}

// timing experiments (garbage results): drain the accumulator without scanning it
__device__ __forceinline__ void epilogue_loads_only(uint32_t t_addr, float& best) {
  uint32_t ra[32];
  uint32_t x = 0;
#pragma unroll 1
  for (int c = 0; c < COLS_PER_WARP / 32; ++c) {
    tc_ld32(t_addr + c * 32, ra);
    tc_wait_ld();
    x ^= ra[0] ^ ra[31];
  }
  if (x == 0x12345678u) best = 0.f;
}

__device__ __forceinline__ void epilogue_tile(bool top1, uint32_t t_addr, int col_base, int m, int etid, float* ring_v, int* ring_i,
                                              float& best, float& second, float& thr, int& cnt, bool& ovf, int experiment = 0) {
  if (experiment == 2) { epilogue_loads_only(t_addr, best); return; }
  if (experiment == 3) return;
  if (experiment == 4) { epilogue_max_only(t_addr, best); return; }
  if (top1)
    epilogue_half<true>(t_addr, col_base, m, etid, ring_v, ring_i, best, second, thr, cnt, ovf);
  else
    epilogue_half<false>(t_addr, col_base, m, etid, ring_v, ring_i, best, second, thr, cnt, ovf);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
    match_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base, sB = base + STAGES * A_STAGE_BYTES;
  float* ring_v = reinterpret_cast<float*>(smem + SMEM_RING_V);
  int* ring_i = reinterpret_cast<int*>(smem + SMEM_RING_I);
  const uint32_t bars = base + SMEM_BARS;
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tfull0 = bars + 16 * STAGES, tempty0 = tfull0 + 8 * NBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SMEM_BARS + 16 * STAGES + 16 * NBUF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_rows = tc_rows(P);
  const long long total_tiles = (long long)((n_rows + TBM - 1) / TBM) * P.col_tiles;
  // a device-side row count can leave fewer tiles than CTAs: the first `g_eff` CTAs then own exactly one tile each, the
  // others none, so that spans are never empty in the middle of a row block (the slot arithmetic relies on it)
  const long long g_eff = min((long long)gridDim.x, total_tiles);
  const bool has_work = (long long)blockIdx.x < g_eff;
  const long long t_begin = has_work ? (total_tiles * blockIdx.x) / g_eff : 0;
  const long long t_end = has_work ? (total_tiles * (blockIdx.x + 1)) / g_eff : 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512u);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (whole warp walks the loop, one elected lane issues) =====
    {
      uint32_t stage = 0, phase = 0;
      for (long long t = t_begin; t < t_end; ++t) {
        const int rb = (int)(t / P.col_tiles), ct = (int)(t % P.col_tiles);
#pragma unroll 1
        for (int kb = 0; kb < P.kb; ++kb) {
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          if (elect_one()) {
            if (P.experiment && (t > t_begin || kb >= STAGES)) {
              mbar_arrive(full0 + 8 * stage);   // no data movement: MMAs re-read stale shared memory
            } else {
              mbar_expect_tx(full0 + 8 * stage, A_STAGE_BYTES + B_STAGE_BYTES);
              tma_load_2d(sA + stage * A_STAGE_BYTES, &map_a, full0 + 8 * stage, kb * TBK, rb * TBM);
              tma_load_2d(sB + stage * B_STAGE_BYTES, &map_b, full0 + 8 * stage, kb * TBK, ct * TBN);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp-uniform loop, one elected lane issues) =====
    {
      uint32_t stage = 0, phase = 0;
      long long it = 0;
      const uint64_t da0 = umma_desc_k_sw128(sA), db0 = umma_desc_k_sw128(sB);
      long long w_tempty = 0, w_full = 0, t_start = clock64();
      for (long long t = t_begin; t < t_end; ++t, ++it) {
        const uint32_t buf = (uint32_t)(it % NBUF);
        long long c0 = clock64();
        mbar_wait(tempty0 + 8 * buf, (uint32_t)((it / NBUF) & 1) ^ 1);
        w_tempty += clock64() - c0;
        if (P.dbg && blockIdx.x == 0 && it >= 8 && it < 24 && lane == 0) P.dbg[1200 + (it - 8) * 16 + 0] = clock64();  // tempty acquired
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * TBN;
#pragma unroll 1
        for (int kb = 0; kb < P.kb; ++kb) {
          c0 = clock64();
          mbar_wait(full0 + 8 * stage, phase);
          w_full += clock64() - c0;
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = da0 + (uint64_t)(stage * (A_STAGE_BYTES >> 4));
            const uint64_t db = db0 + (uint64_t)(stage * (B_STAGE_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < TBK / UMMA_K; ++k) {
              // advancing 16 fp16 = 32 B inside the 128 B swizzle row: +2 in the (>>4) start-address field
              tc_mma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC, (kb | k) != 0 ? 1u : 0u);
            }
            tc_commit(empty0 + 8 * stage);  // frees the smem stage once these MMAs have read it
            if (kb == P.kb - 1) tc_commit(tfull0 + 8 * buf);  // accumulator complete -> epilogue
          }
          __syncwarp();
          if (kb == P.kb - 1 && P.dbg && blockIdx.x == 0 && it >= 8 && it < 24 && lane == 0) P.dbg[1200 + (it - 8) * 16 + 1] = clock64();  // commit issued
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (P.dbg && lane == 0) {
        P.dbg[blockIdx.x * 8 + 0] = clock64() - t_start;
        P.dbg[blockIdx.x * 8 + 1] = w_tempty;
        P.dbg[blockIdx.x * 8 + 2] = w_full;
        P.dbg[blockIdx.x * 8 + 3] = it;
      }
    }
  } else {
    // ===== epilogue: warps 2..9, TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int tid = q * 32 + lane;        // row inside the block == TMEM lane
    const int etid = half * 128 + tid;    // index among the epilogue threads
    int cur_rb = -1;
    float best = -INFINITY, second = -INFINITY, thr = -INFINITY;
    int cnt = 0;
    bool ovf = false, active = false;
    long long it = 0;
    long long w_tfull = 0, e_start = clock64();
    for (long long t = t_begin; t < t_end; ++t, ++it) {
      const int rb = (int)(t / P.col_tiles), ct = (int)(t % P.col_tiles);
      if (rb != cur_rb) {
        if (cur_rb >= 0) flush_slot(P, n_rows, total_tiles, cur_rb, tid, half, ring_v, ring_i, cnt, ovf, best, second);
        cur_rb = rb;
        const int row = rb * TBM + tid;
        active = (row < n_rows) && (P.nz[row] != 0);
        best = second = -INFINITY;
        thr = active ? ((P.seed && P.top1) ? __ldg(P.seed + row) - MARGIN : -INFINITY) : INFINITY;
        cnt = 0;
        ovf = false;
      }
      const uint32_t buf = (uint32_t)(it % NBUF);
      const long long c0 = clock64();
      mbar_wait(tfull0 + 8 * buf, (uint32_t)((it / NBUF) & 1));
      w_tfull += clock64() - c0;
      if (P.dbg && blockIdx.x == 0 && it >= 8 && it < 24 && lane == 0) P.dbg[1200 + (it - 8) * 16 + 2 + (warp - 2) * 2] = clock64();  // tfull seen
      tc_fence_after();
      epilogue_tile(P.top1 != 0, tmem_base + ((uint32_t)(q * 32) << 16) + buf * TBN + half * COLS_PER_WARP, ct * TBN + half * COLS_PER_WARP,
                    P.m, etid, ring_v, ring_i, best, second, thr, cnt, ovf);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
      if (P.dbg && blockIdx.x == 0 && it >= 8 && it < 24 && lane == 0) P.dbg[1200 + (it - 8) * 16 + 3 + (warp - 2) * 2] = clock64();  // released
    }
    if (P.dbg && warp == 2 && lane == 0) {
      P.dbg[blockIdx.x * 8 + 4] = clock64() - e_start;
      P.dbg[blockIdx.x * 8 + 5] = w_tfull;
    }
    if (cur_rb >= 0) flush_slot(P, n_rows, total_tiles, cur_rb, tid, half, ring_v, ring_i, cnt, ovf, best, second);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Version 2 of the candidate search for dp <= 384: CTA PAIRS (thread-block cluster of 2).
//   * each CTA keeps its own 128-row query block RESIDENT in shared memory (dp/64 chunks of 16 KB, reloaded only when
//     the row block changes), so only the database operand is streamed;
//   * the two CTAs of a pair work on the same 256-column database tile at the same time; each TMA-loads one half of the
//     tile (128 rows) and MULTICASTS it into both CTAs' shared memory, so every database byte crosses L2->SM once per pair;
//   * L2->SM traffic per CTA and 64-wide K chunk drops from (128 + 256) x 128 B to 128 x 128 B.
// Stage release needs both consumers: the MMA warps commit with a multicast arrive onto both CTAs' `empty` barriers
// (count 2).  Everything after the MMA (TMEM double buffering, epilogue, candidate lists, slots) is as in version 1, with
// "cluster" in place of "CTA" for the span / slot arithmetic.
constexpr int STAGES2 = (TBN == 256) ? 3 : 5;
constexpr int A_MAX_KB = 6;  // resident query block: up to 384 columns
constexpr uint32_t B_HALF_BYTES = B_STAGE_BYTES / 2;
constexpr uint32_t S2_A = 0;
constexpr uint32_t S2_B = A_MAX_KB * A_STAGE_BYTES;
constexpr uint32_t S2_RING_V = S2_B + STAGES2 * B_STAGE_BYTES;
constexpr uint32_t S2_RING_I = S2_RING_V + EPI_THREADS * CAP * 4;
constexpr uint32_t S2_BARS = S2_RING_I + EPI_THREADS * CAP * 4;
constexpr uint32_t S2_TOTAL = S2_BARS + 256 + 1024;

__device__ __forceinline__ void flush_slot2(const TcParams& P, int n_rows, int rb, int tid, int half, float* ring_v, int* ring_i,
                                            int cnt, bool ovf, float best, float second, long long total_units, int units_per_rp) {
  const int etid = half * 128 + tid;
  const int row = rb * TBM + tid;
  if (row >= n_rows) return;
  const long long t0 = (long long)(rb >> 1) * units_per_rp;   // first unit of this row-block pair
  const long long g = min((long long)(gridDim.x >> 1), total_units);   // clusters that own a (non-empty) span
  long long c0 = (t0 * g) / total_units;
  while ((total_units * (c0 + 1)) / g <= t0) ++c0;
  while (c0 > 0 && (total_units * c0) / g > t0) --c0;
  const int slot = (int)((blockIdx.x >> 1) - c0) * HALVES + half;
  const long long o = ((long long)row * P.slots + slot);
  P.cand_n[o] = cnt | (ovf ? (1 << 30) : 0);
  P.slot_top2[o] = make_float2(best, second);
  for (int e = 0; e < cnt; ++e) {
    P.cand_v[o * CAP + e] = ring_v[e * EPI_THREADS + etid];
    P.cand_i[o * CAP + e] = ring_i[e * EPI_THREADS + etid];
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
    match_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base + S2_A, sB = base + S2_B;
  float* ring_v = reinterpret_cast<float*>(smem + S2_RING_V);
  int* ring_i = reinterpret_cast<int*>(smem + S2_RING_I);
  const uint32_t bars = base + S2_BARS;
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES2, tfull0 = bars + 16 * STAGES2, tempty0 = tfull0 + 8 * NBUF;
  const uint32_t afull = tempty0 + 8 * NBUF, aempty = afull + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S2_BARS + 16 * STAGES2 + 16 * NBUF + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_cta_rank();
  const long long clusters = gridDim.x >> 1, cid = blockIdx.x >> 1;
  const int n_rows = tc_rows(P);
  const int row_pairs = (n_rows + 2 * TBM - 1) / (2 * TBM);
  const long long total_units = (long long)row_pairs * P.col_tiles;
  // fewer units than clusters (device-side row count): the first `c_eff` clusters own one unit each, the others none
  const long long c_eff = min(clusters, total_units);
  const long long u_begin = cid < c_eff ? (total_units * cid) / c_eff : 0;
  const long long u_end = cid < c_eff ? (total_units * (cid + 1)) / c_eff : 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 2);   // both CTAs of the pair read every stage
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, EPI_WARPS);
    }
    mbar_init(afull, 1);
    mbar_init(aempty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers must be initialised before any multicast can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (whole warp walks the loop, one elected lane issues) =====
    {
      uint32_t stage = 0, phase = 0, aphase = 0;
      int cur_rp = -1;
      for (long long u = u_begin; u < u_end; ++u) {
        const int rp = (int)(u / P.col_tiles), ct = (int)(u % P.col_tiles);
        if (rp != cur_rp) {   // new resident query block
          cur_rp = rp;
          mbar_wait(aempty, aphase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(afull, (uint32_t)P.kb * A_STAGE_BYTES);
            for (int kb = 0; kb < P.kb; ++kb)
              tma_load_2d(sA + kb * A_STAGE_BYTES, &map_a, afull, kb * TBK, (2 * rp + (int)rank) * TBM);
          }
          __syncwarp();
          aphase ^= 1;
        }
#pragma unroll 1
        for (int kb = 0; kb < P.kb; ++kb) {
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(full0 + 8 * stage, B_STAGE_BYTES);   // my half + the peer's half
            tma_load_2d_mc(sB + stage * B_STAGE_BYTES + rank * B_HALF_BYTES, &map_b, full0 + 8 * stage, kb * TBK,
                           ct * TBN + (int)rank * (TBN / 2), (uint16_t)3);
          }
          __syncwarp();
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp-uniform loop, one elected lane issues) =====
    {
      uint32_t stage = 0, phase = 0, aphase = 0;
      long long it = 0;
      int cur_rp = -1;
      const uint64_t da0 = umma_desc_k_sw128(sA), db0 = umma_desc_k_sw128(sB);
      for (long long u = u_begin; u < u_end; ++u, ++it) {
        const int rp = (int)(u / P.col_tiles);
        const uint32_t buf = (uint32_t)(it % NBUF);
        mbar_wait(tempty0 + 8 * buf, (uint32_t)((it / NBUF) & 1) ^ 1);
        if (rp != cur_rp) {
          cur_rp = rp;
          mbar_wait(afull, aphase);
          aphase ^= 1;
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * TBN;
        const bool last_of_rp = (u + 1 == u_end) || ((int)((u + 1) / P.col_tiles) != rp);
#pragma unroll 1
        for (int kb = 0; kb < P.kb; ++kb) {
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = da0 + (uint64_t)(kb * (A_STAGE_BYTES >> 4));
            const uint64_t db = db0 + (uint64_t)(stage * (B_STAGE_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < TBK / UMMA_K; ++k)
              tc_mma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC, (kb | k) != 0 ? 1u : 0u);
            tc_commit_mc(empty0 + 8 * stage, (uint16_t)3);   // frees the stage in both CTAs once these MMAs have read it
            if (kb == P.kb - 1) {
              tc_commit(tfull0 + 8 * buf);
              if (last_of_rp) tc_commit(aempty);   // resident block may be overwritten after these MMAs
            }
          }
          __syncwarp();
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===== epilogue (as in version 1) =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int tid = q * 32 + lane;
    const int etid = half * 128 + tid;
    int cur_rb = -1;
    float best = -INFINITY, second = -INFINITY, thr = -INFINITY;
    int cnt = 0;
    bool ovf = false, active = false;
    long long it = 0;
    for (long long u = u_begin; u < u_end; ++u, ++it) {
      const int rp = (int)(u / P.col_tiles), ct = (int)(u % P.col_tiles);
      const int rb = 2 * rp + (int)rank;
      if (rb != cur_rb) {
        if (cur_rb >= 0) flush_slot2(P, n_rows, cur_rb, tid, half, ring_v, ring_i, cnt, ovf, best, second, total_units, P.col_tiles);
        cur_rb = rb;
        const int row = rb * TBM + tid;
        active = (row < n_rows) && (P.nz[row] != 0);
        best = second = -INFINITY;
        thr = active ? ((P.seed && P.top1) ? __ldg(P.seed + row) - MARGIN : -INFINITY) : INFINITY;
        cnt = 0;
        ovf = false;
      }
      const uint32_t buf = (uint32_t)(it % NBUF);
      mbar_wait(tfull0 + 8 * buf, (uint32_t)((it / NBUF) & 1));
      tc_fence_after();
      epilogue_tile(P.top1 != 0, tmem_base + ((uint32_t)(q * 32) << 16) + buf * TBN + half * COLS_PER_WARP, ct * TBN + half * COLS_PER_WARP,
                    P.m, etid, ring_v, ring_i, best, second, thr, cnt, ovf);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
    }
    if (cur_rb >= 0) flush_slot2(P, n_rows, cur_rb, tid, half, ring_v, ring_i, cnt, ovf, best, second, total_units, P.col_tiles);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA may exit while its peer can still multicast into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Version 3: CTA-pair MMA (tcgen05.mma.cta_group::2).  One instruction issued by the leader CTA computes a 256 x 256 x 16
// block on the tensor cores of both SMs of the pair: each CTA holds its own 128 query rows (A) and HALF of the 256-column
// database tile (B) in its own shared memory, and receives the 128 x 256 accumulator rows of its query block in its own
// TMEM.  Against version 2 (two independent M128 N256 MMAs fed by multicast) the shared-memory operand traffic per MMA
// drops from 12 KB to 8 KB per SM -- the measured limiter of versions 1/2 (profiles/r1_kernel_bench.txt) -- and every
// database byte still crosses L2 -> SM once per pair without multicast.
//   RESIDENT (dp <= 384): the query block stays in shared memory for the whole row of column tiles; only B is streamed.
//   otherwise (dp up to 1024): A and B chunks are streamed together.
// Barriers: `full` / `afull` / `tempty` live in the leader (both CTAs' TMA loads credit the leader's barriers through
// .cta_group::2 loads; the peer's epilogue warps arrive remotely); `empty` / `aempty` / `tfull` exist in both CTAs and are
// signalled by multicast tcgen05.commit.
constexpr int STAGES3 = 6;
constexpr uint32_t S3_A = 0;                                   // RESIDENT: dp/64 chunks; streaming: STAGES3 chunks
constexpr uint32_t S3_B = A_MAX_KB * A_STAGE_BYTES;            // STAGES3 x (128 database rows x 64 k) = 16 KB each
constexpr uint32_t S3_RING_V = S3_B + STAGES3 * B_HALF_BYTES;
constexpr uint32_t S3_RING_I = S3_RING_V + EPI_THREADS * CAP * 4;
constexpr uint32_t S3_BARS = S3_RING_I + EPI_THREADS * CAP * 4;
constexpr uint32_t S3_TOTAL = S3_BARS + 256 + 1024;
constexpr uint32_t IDESC3 = umma_idesc_f16(2 * TBM, TBN, 0);
static_assert(TBN == 256 && NBUF == 2, "version 3 is written for 256-column tiles");
static_assert(STAGES3 <= A_MAX_KB, "streamed A chunks reuse the resident region");

template <bool RESIDENT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
    match_tc3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base + S3_A, sB = base + S3_B;
  float* ring_v = reinterpret_cast<float*>(smem + S3_RING_V);
  int* ring_i = reinterpret_cast<int*>(smem + S3_RING_I);
  const uint32_t bars = base + S3_BARS;
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES3, tfull0 = bars + 16 * STAGES3, tempty0 = tfull0 + 8 * NBUF;
  const uint32_t afull = tempty0 + 8 * NBUF, aempty = afull + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S3_BARS + 16 * STAGES3 + 16 * NBUF + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_cta_rank();
  const long long clusters = gridDim.x >> 1, cid = blockIdx.x >> 1;
  const int n_rows = tc_rows(P);
  const int row_pairs = (n_rows + 2 * TBM - 1) / (2 * TBM);
  const long long total_units = (long long)row_pairs * P.col_tiles;
  const long long c_eff = min(clusters, total_units);
  const long long u_begin = cid < c_eff ? (total_units * cid) / c_eff : 0;
  const long long u_end = cid < c_eff ? (total_units * (cid + 1)) / c_eff : 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < STAGES3; ++s) {
      mbar_init(full0 + 8 * s, 1);    // leader: one arrive.expect_tx per phase, bytes from both CTAs' loads
      mbar_init(empty0 + 8 * s, 1);   // both CTAs: multicast commit from the leader's MMA warp
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(tfull0 + 8 * b, 1);                // both CTAs: multicast commit
      mbar_init(tempty0 + 8 * b, 2 * EPI_WARPS);   // leader: epilogue warps of both CTAs
    }
    mbar_init(afull, 1);
    mbar_init(aempty, 1);
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
    // ===== TMA producer: each CTA loads its own query rows and its half of the database tile =====
    uint32_t stage = 0, phase = 0, aphase = 0;
    int cur_rp = -1;
    const uint32_t afull_leader = mapa_cluster(afull, 0);
    const bool dbg = P.dbg != nullptr;   // pipeline counters only when asked for (VFMREG_TC_DEBUG=1)
    long long w_empty = 0;
    const long long p_start = dbg ? clock64() : 0;
    for (long long u = u_begin; u < u_end; ++u) {
      const int rp = (int)(u / P.col_tiles), ct = (int)(u % P.col_tiles);
      if (RESIDENT && rp != cur_rp) {
        cur_rp = rp;
        mbar_wait(aempty, aphase ^ 1);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(afull, 2u * (uint32_t)P.kb * A_STAGE_BYTES);
          for (int kb = 0; kb < P.kb; ++kb)
            tma_load_2d_2sm(sA + kb * A_STAGE_BYTES, &map_a, afull_leader, kb * TBK, (2 * rp + (int)rank) * TBM);
        }
        __syncwarp();
        aphase ^= 1;
      }
#pragma unroll 1
      for (int kb = 0; kb < P.kb; ++kb) {
        const long long c0 = dbg ? clock64() : 0;
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        if (dbg) w_empty += clock64() - c0;
        if (elect_one()) {
          const uint32_t full_leader = mapa_cluster(full0 + 8 * stage, 0);
          if (rank == 0) mbar_expect_tx(full0 + 8 * stage, RESIDENT ? 2u * B_HALF_BYTES : 2u * (B_HALF_BYTES + A_STAGE_BYTES));
          if (!RESIDENT) tma_load_2d_2sm(sA + stage * A_STAGE_BYTES, &map_a, full_leader, kb * TBK, (2 * rp + (int)rank) * TBM);
          tma_load_2d_2sm(sB + stage * B_HALF_BYTES, &map_b, full_leader, kb * TBK, ct * TBN + (int)rank * (TBN / 2));
        }
        __syncwarp();
        if (++stage == STAGES3) { stage = 0; phase ^= 1; }
      }
    }
    if (dbg && lane == 0) {
      P.dbg[blockIdx.x * 8 + 6] = w_empty;
      P.dbg[blockIdx.x * 8 + 7] = clock64() - p_start;
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the leader CTA only =====
    if (rank == 0) {
      uint32_t stage = 0, phase = 0, aphase = 0;
      long long it = 0;
      int cur_rp = -1;
      const uint64_t da0 = umma_desc_k_sw128(sA), db0 = umma_desc_k_sw128(sB);
      const bool dbg = P.dbg != nullptr;
      long long w_tempty = 0, w_full = 0;
      const long long t_start = dbg ? clock64() : 0;
      for (long long u = u_begin; u < u_end; ++u, ++it) {
        const int rp = (int)(u / P.col_tiles);
        const uint32_t buf = (uint32_t)(it % NBUF);
        long long c0 = dbg ? clock64() : 0;
        mbar_wait(tempty0 + 8 * buf, (uint32_t)((it / NBUF) & 1) ^ 1);
        if (dbg) w_tempty += clock64() - c0;
        if (RESIDENT && rp != cur_rp) {
          cur_rp = rp;
          mbar_wait(afull, aphase);
          aphase ^= 1;
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * TBN;
        const bool last_of_rp = (u + 1 == u_end) || ((int)((u + 1) / P.col_tiles) != rp);
#pragma unroll 1
        for (int kb = 0; kb < P.kb; ++kb) {
          if (dbg) c0 = clock64();
          mbar_wait(full0 + 8 * stage, phase);
          if (dbg) w_full += clock64() - c0;
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = da0 + (uint64_t)((RESIDENT ? kb : (int)stage) * (A_STAGE_BYTES >> 4));
            const uint64_t db = db0 + (uint64_t)(stage * (B_HALF_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < TBK / UMMA_K; ++k)
              tc_mma_f16_2sm(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC3, (kb | k) != 0 ? 1u : 0u);
            tc_commit_2sm(empty0 + 8 * stage, (uint16_t)3);   // stage free in both CTAs once these MMAs have read it
            if (kb == P.kb - 1) {
              tc_commit_2sm(tfull0 + 8 * buf, (uint16_t)3);   // accumulator halves complete -> both epilogues
              if (RESIDENT && last_of_rp) tc_commit_2sm(aempty, (uint16_t)3);
            }
          }
          __syncwarp();
          if (++stage == STAGES3) { stage = 0; phase ^= 1; }
        }
      }
      if (dbg && lane == 0) {
        P.dbg[blockIdx.x * 8 + 0] = clock64() - t_start;
        P.dbg[blockIdx.x * 8 + 1] = w_tempty;
        P.dbg[blockIdx.x * 8 + 2] = w_full;
        P.dbg[blockIdx.x * 8 + 3] = it;
      }
    }
  } else {
    // ===== epilogue (as in versions 1/2): this CTA's 128 rows x 256 columns =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int tid = q * 32 + lane;
    const int etid = half * 128 + tid;
    int cur_rb = -1;
    float best = -INFINITY, second = -INFINITY, thr = -INFINITY;
    int cnt = 0;
    bool ovf = false, active = false;
    long long it = 0;
    const bool dbg = P.dbg != nullptr;
    long long w_tfull = 0;
    const long long e_start = dbg ? clock64() : 0;
    for (long long u = u_begin; u < u_end; ++u, ++it) {
      const int rp = (int)(u / P.col_tiles), ct = (int)(u % P.col_tiles);
      const int rb = 2 * rp + (int)rank;
      if (rb != cur_rb) {
        if (cur_rb >= 0) flush_slot2(P, n_rows, cur_rb, tid, half, ring_v, ring_i, cnt, ovf, best, second, total_units, P.col_tiles);
        cur_rb = rb;
        const int row = rb * TBM + tid;
        active = (row < n_rows) && (P.nz[row] != 0);
        best = second = -INFINITY;
        thr = active ? ((P.seed && P.top1) ? __ldg(P.seed + row) - MARGIN : -INFINITY) : INFINITY;
        cnt = 0;
        ovf = false;
      }
      const uint32_t buf = (uint32_t)(it % NBUF);
      const long long c0 = dbg ? clock64() : 0;
      mbar_wait(tfull0 + 8 * buf, (uint32_t)((it / NBUF) & 1));
      if (dbg) w_tfull += clock64() - c0;
      tc_fence_after();
      epilogue_tile(P.top1 != 0, tmem_base + ((uint32_t)(q * 32) << 16) + buf * TBN + half * COLS_PER_WARP, ct * TBN + half * COLS_PER_WARP,
                    P.m, etid, ring_v, ring_i, best, second, thr, cnt, ovf, P.experiment);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_cluster(tempty0 + 8 * buf, 0));
    }
    if (dbg && warp == 2 && lane == 0) {
      P.dbg[blockIdx.x * 8 + 4] = clock64() - e_start;
      P.dbg[blockIdx.x * 8 + 5] = w_tfull;
    }
    if (cur_rb >= 0) flush_slot2(P, n_rows, cur_rb, tid, half, ring_v, ring_i, cnt, ovf, best, second, total_units, P.col_tiles);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs are done with the pair's TMEM and with each other's barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512u);
  }
}

// canonical fp32 inner product: fmaf chain over k ascending (dp % 32 == 0 here, rows 16-byte aligned).  The chain is
// sequential by definition; the loads are not: 8 float4 of each row are requested before the 32 dependent fmaf run.
__device__ __forceinline__ float canon_dot(const float* __restrict__ x, const float* __restrict__ y, int dp) {
  float acc = 0.0f;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const float4* y4 = reinterpret_cast<const float4*>(y);
#pragma unroll 1
  for (int k = 0; k < dp / 4; k += 8) {
    float4 a[8], b[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a[u] = __ldg(x4 + k + u);
      b[u] = __ldg(y4 + k + u);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      acc = fmaf(a[u].x, b[u].x, acc);
      acc = fmaf(a[u].y, b[u].y, acc);
      acc = fmaf(a[u].z, b[u].z, acc);
      acc = fmaf(a[u].w, b[u].w, acc);
    }
  }
  return acc;
}

// Re-rank, step 1a: one thread per (query row, slot) rebuilds the row's recording threshold from the per-slot approximate
// top-2 (second largest of all slot bests / runner-ups, minus the margin), marks the candidates below it -inf and appends
// the survivors (typically 1-3 per row) to a compact work list (one atomic per warp).
__global__ void __launch_bounds__(128)
    rerank_select_kernel(int n, const int* __restrict__ n_dev, int slots, int top1, const uint8_t* __restrict__ nz, float* __restrict__ cand_v,
                         const int* __restrict__ cand_n, const float2* __restrict__ slot_top2, int* __restrict__ work,
                         int* __restrict__ work_count, uint8_t* __restrict__ row_overflow) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  if (n_dev) n = min(n, __ldg(n_dev));
  int keep[CAP];
  int n_keep = 0;
  if (g < (long long)n * slots) {
    const int cnt = cand_n[g] & 0xFFFF;
    const int row = (int)(g / slots);
    if (cnt > 0 && nz[row]) {
      float t1 = -INFINITY, t2 = -INFINITY;
      bool overflow = false;
      for (int s = 0; s < slots; ++s) {
        const int c = cand_n[(long long)row * slots + s];
        if ((c & 0xFFFF) == 0 && !((c >> 30) & 1)) continue;
        overflow |= (c >> 30) & 1;
        const float2 t = slot_top2[(long long)row * slots + s];
        if (t.x > t1) { t2 = fmaxf(t1, t.y); t1 = t.x; } else { t2 = fmaxf(t2, t.x); }
      }
      if (overflow && row_overflow) row_overflow[row] = 1;
      if (!overflow) {  // overflowed rows are redone exactly
        const float thr = (top1 ? t1 : t2) - MARGIN;
#pragma unroll
        for (int e = 0; e < CAP; ++e) {
          if (e < cnt) {
            if (cand_v[g * CAP + e] >= thr) keep[n_keep++] = e;
            else cand_v[g * CAP + e] = -INFINITY;
          }
        }
      }
    }
  }
  // warp-aggregated append
  int incl = n_keep;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  int base = 0;
  if (lane == 31 && total > 0) base = atomicAdd(work_count, total);
  base = __shfl_sync(0xffffffffu, base, 31);
  int o = base + incl - n_keep;
#pragma unroll
  for (int e = 0; e < CAP; ++e)
    if (e < n_keep) work[o + e] = (int)(g * CAP + keep[e]);
}

// (score, index) packed so that an unsigned 64-bit max picks the largest score and, on equal scores, the lowest index.
// Canonical scores are never -0 (the fmaf chain starts from +0), so the bit order of the transformed float is the float order.
__device__ __forceinline__ unsigned long long pack_best(float score, int idx) {
  uint32_t u = __float_as_uint(score);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)idx);
}
__device__ __forceinline__ void unpack_best(unsigned long long key, float& score, int& idx) {
  uint32_t u = (uint32_t)(key >> 32);
  u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
  score = __uint_as_float(u);
  idx = (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFu));
}

// Re-rank, step 1b: L = 16 lanes per surviving candidate.  Lane l owns the l-th contiguous segment of the two rows (dp / 16
// elements = dp / 64 float4 per row, at most PER_MAX), so every load of a candidate is in flight at once; the canonical fmaf chain over k
// ascending is then walked lane by lane, the accumulator handed on by shuffle -- the same operation order as one thread
// running the whole chain.  The exact score replaces the approximate one; in top-1 mode it is also folded into the
// row's packed (score, index) maximum, which makes the per-row pick a single atomic.
template <int PER_MAX>
__global__ void __launch_bounds__(128)
    rerank_dot_kernel(const float* __restrict__ a, const float* __restrict__ b, int dp, int slots, float* __restrict__ cand_v,
                      const int* __restrict__ cand_i, const int* __restrict__ work, const int* __restrict__ work_count,
                      unsigned long long* __restrict__ row_key) {
  constexpr int L = 16;
  const int count = *work_count;
  const int lane = threadIdx.x & 31, gl = lane & (L - 1), gbase = lane & ~(L - 1);
  const int per = dp / (4 * L);   // float4 per lane and row (dp % 64 == 0, dp <= 64 PER_MAX)
  const int groups = (gridDim.x * blockDim.x) / L;
  const int rounds = (count + groups - 1) / groups;
  int w = (blockIdx.x * blockDim.x + threadIdx.x) / L;
  for (int r = 0; r < rounds; ++r, w += groups) {   // every lane walks the loop: the shuffles need the whole warp
    const bool live = w < count;
    int ent = 0, row = 0, col = 0;
    float4 av[PER_MAX], bv[PER_MAX];
    if (live) {
      ent = work[w];
      row = ent / (slots * CAP);
      col = cand_i[ent];
      const float4* a4 = reinterpret_cast<const float4*>(a + (long long)row * dp) + gl * per;
      const float4* b4 = reinterpret_cast<const float4*>(b + (long long)col * dp) + gl * per;
#pragma unroll
      for (int u = 0; u < PER_MAX; ++u)
        if (u < per) {
          av[u] = __ldg(a4 + u);
          bv[u] = __ldg(b4 + u);
        }
    }
    float acc = 0.0f;
#pragma unroll 1
    for (int sl = 0; sl < L; ++sl) {
      if (live && gl == sl) {
#pragma unroll
        for (int u = 0; u < PER_MAX; ++u)
          if (u < per) {
            acc = fmaf(av[u].x, bv[u].x, acc);
            acc = fmaf(av[u].y, bv[u].y, acc);
            acc = fmaf(av[u].z, bv[u].z, acc);
            acc = fmaf(av[u].w, bv[u].w, acc);
          }
      }
      acc = __shfl_sync(0xffffffffu, acc, gbase + sl);   // lane sl's running sum -> every lane of the group
    }
    if (live && gl == 0) {
      cand_v[ent] = acc;
      if (row_key) atomicMax(row_key + row, pack_best(acc, col));
    }
  }
}

// Re-rank, step 2 in top-1 mode: unpack the row's (score, index) maximum; overflowed rows go to exact_rows_kernel.
__global__ void __launch_bounds__(128)
    rerank_finish_kernel(int n, const int* __restrict__ n_dev, const uint8_t* __restrict__ nz, const unsigned long long* __restrict__ row_key,
                         const uint8_t* __restrict__ row_overflow, int32_t* __restrict__ idx, float* __restrict__ best,
                         int* __restrict__ redo_list, int* __restrict__ redo_count) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, __ldg(n_dev));
  if (row >= n) return;
  if (!nz[row]) {  // all-zero query: every canonical inner product is exactly +0 -> lowest index wins
    idx[row] = 0;
    if (best) best[row] = 0.0f;
    return;
  }
  const unsigned long long key = row_key[row];
  if (row_overflow[row] || key == 0ULL) {
    redo_list[atomicAdd(redo_count, 1)] = row;
    return;
  }
  float sc;
  int j;
  unpack_best(key, sc, j);
  idx[row] = j;
  if (best) best[row] = sc;
}

// Re-rank, step 2: one thread per query row picks the exact top-2 (lowest index on ties); rows whose list overflowed are
// queued for exact_rows_kernel.
__global__ void __launch_bounds__(128)
    rerank_pick_kernel(int n, const int* __restrict__ n_dev, int m, int slots, const uint8_t* __restrict__ nz,
                       const float* __restrict__ cand_v, const int* __restrict__ cand_i, const int* __restrict__ cand_n,
                       int32_t* __restrict__ idx, float* __restrict__ best, float* __restrict__ sec, int* __restrict__ redo_list,
                       int* __restrict__ redo_count) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, __ldg(n_dev));
  if (row >= n) return;
  if (!nz[row]) {  // all-zero query: every canonical inner product is exactly +0 -> lowest index wins
    idx[row] = 0;
    if (best) best[row] = 0.0f;
    if (sec) sec[row] = (m > 1) ? 0.0f : -INFINITY;
    return;
  }
  float b1 = -INFINITY, b2 = -INFINITY;
  int bi = 0x7fffffff;
  bool overflow = false;
  for (int s = 0; s < slots; ++s) {
    const long long o = (long long)row * slots + s;
    const int cn = cand_n[o];
    overflow |= (cn >> 30) & 1;
    const int c = cn & 0xFFFF;
    for (int e = 0; e < c; ++e) {
      const float x = cand_v[o * CAP + e];
      if (x == -INFINITY) continue;
      const int j = cand_i[o * CAP + e];
      if (x > b1 || (x == b1 && j < bi)) {
        b2 = b1;
        b1 = x;
        bi = j;
      } else if (x > b2) {
        b2 = x;
      }
    }
  }
  if (overflow) {
    redo_list[atomicAdd(redo_count, 1)] = row;
    return;
  }
  idx[row] = (bi == 0x7fffffff) ? 0 : bi;
  if (best) best[row] = b1;
  if (sec) sec[row] = b2;
}

// Exact scan of all columns for the (rare) rows whose candidate list overflowed: one CTA per listed row.
__global__ void __launch_bounds__(256)
    exact_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, int m, int dp, const int* __restrict__ redo_list,
                      const int* __restrict__ redo_count, int32_t* __restrict__ idx, float* __restrict__ best,
                      float* __restrict__ sec) {
  __shared__ float s1[256], s2[256];
  __shared__ int si[256];
  const int count = *redo_count;
  for (int li = blockIdx.x; li < count; li += gridDim.x) {
    const int row = redo_list[li];
    const float* ar = a + (long long)row * dp;
    float b1 = -INFINITY, b2 = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = threadIdx.x; j < m; j += 256) {
      const float x = canon_dot(ar, b + (long long)j * dp, dp);
      if (x > b1) {  // j ascending inside a thread: strict '>' keeps the lowest index
        b2 = b1;
        b1 = x;
        bi = j;
      } else if (x > b2) {
        b2 = x;
      }
    }
    s1[threadIdx.x] = b1;
    s2[threadIdx.x] = b2;
    si[threadIdx.x] = bi;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int t = 1; t < 256; ++t) {
        const float y1 = s1[t], y2 = s2[t];
        const int yi = si[t];
        const bool y_wins = (y1 > b1) || (y1 == b1 && yi < bi);
        const float lo = y_wins ? b1 : y1;
        b2 = fmaxf(fmaxf(b2, y2), lo);
        if (y_wins) {
          b1 = y1;
          bi = yi;
        }
      }
      idx[row] = (bi == 0x7fffffff) ? 0 : bi;
      if (best) best[row] = b1;
      if (sec) sec[row] = b2;
    }
    __syncthreads();
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_16bit(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t pitch_elems, int box_rows, bool bf16) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return VFMREG_ERR_CUDA;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * 2};
  const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld", (int)r, (long long)rows, (long long)cols);
    return VFMREG_ERR_CUDA;
  }
  return VFMREG_OK;
}

static int make_map_f16(CUtensorMap* map, const void* ptr, int64_t rows, int dp, int box_rows) {
  return make_tmap_16bit(map, ptr, rows, dp, dp, box_rows, false);
}

struct TcPlan {
  int row_blocks, col_tiles, grid, slots;
  long long total;
  bool paired;   // CTA pairs: version 3 (cta_group::2 MMA) or version 2 (resident query block + multicast database tiles)
  int version;   // 1, 2 or 3
};

// VFMREG_MATCH_V1=1 in the environment keeps the single-CTA streaming kernel, VFMREG_MATCH_V2=1 the multicast CTA-pair
// kernel (A/B comparison while tuning)
static bool g_force_v1 = [] { const char* e = getenv("VFMREG_MATCH_V1"); return e && e[0] == '1'; }();
static bool g_force_v2 = [] { const char* e = getenv("VFMREG_MATCH_V2"); return e && e[0] == '1'; }();

// `dynamic`: n is only an upper bound, the kernels read the row count from device memory.  The grid is sized for n; the
// slot bound must then hold for every smaller total: spans of at least one unit cut a row block's col_tiles consecutive
// units into at most col_tiles + 1 pieces (and spans of at most one unit into at most col_tiles).
static TcPlan tc_plan(vfmreg_ctx* ctx, int64_t n, int64_t m, int dp, bool dynamic = false) {
  TcPlan p;
  p.row_blocks = ceil_div(n, TBM);
  p.col_tiles = ceil_div(m, TBN);
  p.version = g_force_v1 ? 1 : ((g_force_v2 && dp <= A_MAX_KB * TBK) ? 2 : (g_force_v2 ? 1 : 3));
  if (p.row_blocks < 2) p.version = 1;
  p.paired = p.version != 1;
  if (p.paired) {
    const int row_pairs = ceil_div(n, 2 * TBM);
    p.total = (long long)row_pairs * p.col_tiles;          // units of (row-block pair, column tile)
    const int clusters = (int)((p.total < ctx->sm_count / 2) ? p.total : ctx->sm_count / 2);
    p.grid = 2 * clusters;
    const long long min_span = p.total / clusters;
    p.slots = dynamic ? p.col_tiles + 1 : (int)((p.col_tiles + min_span - 1) / min_span) + 1;
    if (p.slots > clusters) p.slots = clusters;
    p.slots *= HALVES;   // column groups per span
    return p;
  }
  p.total = (long long)p.row_blocks * p.col_tiles;
  p.grid = (int)((p.total < ctx->sm_count) ? p.total : ctx->sm_count);
  // a row block of col_tiles consecutive tiles is cut by at most ceil(col_tiles / floor(total/grid)) + 1 spans
  const long long min_span = p.total / p.grid;
  p.slots = dynamic ? p.col_tiles + 1 : (int)((p.col_tiles + min_span - 1) / min_span) + 1;
  if (p.slots > p.grid) p.slots = p.grid;
  p.slots *= HALVES;   // column groups per span
  return p;
}

void match_tc_force_v1(bool on) { g_force_v1 = on; }

size_t match_tc_scratch(vfmreg_ctx* ctx, int64_t n, int64_t m, bool dynamic) {
  TcPlan p = tc_plan(ctx, n, m, 64, dynamic);
  const TcPlan p1 = tc_plan(ctx, n, m, 1 << 20, dynamic);
  if (p1.slots > p.slots) p.slots = p1.slots;
  return 2 * arena_bytes((size_t)n * p.slots * CAP, 4) + arena_bytes((size_t)n * p.slots, 4) +
         arena_bytes((size_t)n * p.slots, 8) + arena_bytes((size_t)n + 1, 4) + arena_bytes((size_t)n * p.slots * CAP + 1, 4) +
         arena_bytes((size_t)n, 8) + arena_bytes((size_t)n, 1) + 2048;
}

// a32/b32: renormalised fp32 rows (n x dp), a16/b16: their fp16 copies, nz_a: non-zero flags of the query rows.
// n_dev (optional, device): the number of query rows actually present; n is then the capacity of a16 / a32 / nz_a / idx.
int match_tc(vfmreg_ctx* ctx, const float* a32, const void* a16, const uint8_t* nz_a, int64_t n, const float* b32,
             const void* b16, int64_t m, int dp, int32_t* idx, float* best, float* sec, const int* n_dev, const float* seed) {
  VFM_CHECK_ARG(dp % TBK == 0 && dp <= 1024, "match_tc: padded dim %d must be a multiple of %d and <= 1024 (error bound)", dp, TBK);
  VFM_CHECK_ARG(n > 0 && m > 0 && n < (1LL << 30) && m < (1LL << 30), "match_tc: bad sizes");
  const TcPlan plan = tc_plan(ctx, n, m, dp, n_dev != nullptr);
  float* cand_v = arena_take<float>(ctx, (size_t)n * plan.slots * CAP);
  int* cand_i = arena_take<int>(ctx, (size_t)n * plan.slots * CAP);
  float2* slot_top2 = arena_take<float2>(ctx, (size_t)n * plan.slots);
  int* redo_list = arena_take<int>(ctx, (size_t)n);                             // rows for exact_rows_kernel
  int* work_list = arena_take<int>(ctx, (size_t)n * plan.slots * CAP);          // candidate entries to re-score
  // everything that starts at zero sits in one block: [redo count, work count | row_key | row_overflow | cand_n], one memset
  const size_t z_key = 256, z_ovf = z_key + arena_bytes((size_t)n, 8), z_cn = z_ovf + arena_bytes((size_t)n, 1);
  const size_t z_bytes = z_cn + arena_bytes((size_t)n * plan.slots, 4);
  char* zero = arena_take<char>(ctx, z_bytes);
  if (!cand_v || !cand_i || !slot_top2 || !redo_list || !work_list || !zero) {
    set_error("match_tc: scratch arena too small");
    return VFMREG_ERR_ALLOC;
  }
  int* redo_count = reinterpret_cast<int*>(zero);
  int* work_count = redo_count + 1;
  unsigned long long* row_key = reinterpret_cast<unsigned long long*>(zero + z_key);   // top-1 mode: packed (score, index) maximum
  uint8_t* row_overflow = reinterpret_cast<uint8_t*>(zero + z_ovf);
  int* cand_n = reinterpret_cast<int*>(zero + z_cn);
  VFM_CUDA(cudaMemsetAsync(zero, 0, z_bytes, ctx->stream));
  CUtensorMap map_a, map_b;
  VFM_TRY(make_map_f16(&map_a, a16, n, dp, TBM));
  VFM_TRY(make_map_f16(&map_b, b16, m, dp, plan.paired ? TBN / 2 : TBN));
  TcParams P;
  P.n = (int)n;
  P.m = (int)m;
  P.kb = dp / TBK;
  P.col_tiles = plan.col_tiles;
  P.total_tiles = plan.total;
  P.slots = plan.slots;
  P.nz = nz_a;
  P.cand_v = cand_v;
  P.cand_i = cand_i;
  P.cand_n = cand_n;
  P.slot_top2 = slot_top2;
  P.dbg = nullptr;
  static const int experiment = [] { const char* e = getenv("VFMREG_TC_EXPERIMENT"); return e ? atoi(e) : 0; }();
  P.experiment = experiment;
  static const bool force_top1 = [] { const char* e = getenv("VFMREG_TC_TOP1"); return e && e[0] == '1'; }();  // tuning aid
  P.top1 = (sec == nullptr || force_top1) ? 1 : 0;
  P.n_dev = n_dev;
  P.seed = seed;
  static const bool want_dbg = [] { const char* e = getenv("VFMREG_TC_DEBUG"); return e && e[0] == '1'; }();
  static long long* dbg_dev = nullptr;
  if (want_dbg) {
    if (!dbg_dev) cudaMalloc(&dbg_dev, 2048 * 8 * sizeof(long long));
    cudaMemsetAsync(dbg_dev, 0, 2048 * 8 * sizeof(long long), ctx->stream);
    P.dbg = dbg_dev;
  }
  static bool attr_set = false;
  if (!attr_set) {
    VFM_CUDA(cudaFuncSetAttribute(match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL));
    VFM_CUDA(cudaFuncSetAttribute(match_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S2_TOTAL));
    VFM_CUDA(cudaFuncSetAttribute(match_tc3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S3_TOTAL));
    VFM_CUDA(cudaFuncSetAttribute(match_tc3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S3_TOTAL));
    attr_set = true;
  }
  const int grp = n_dev ? GROUP_MATCH_PRUNED : GROUP_MATCH;
  // Batch mode (vfmreg_register_batch with several lanes): the candidate-search kernels of all lanes go to one of two
  // high-priority streams -- full searches on [0], pruned reverse searches on [1] -- so that they run back to back in pair
  // order and are scheduled ahead of the lanes' small kernels, which fill the SMs beside them.  The lane stream hands
  // over with an event and takes the result back with another.
  cudaStream_t lane = ctx->stream;
  cudaStream_t ks = ctx->match_stream[n_dev ? 1 : 0];
  if (ks) {
    cudaEvent_t ev = ctx->match_ev[ctx->match_ev_head];
    ctx->match_ev_head = (ctx->match_ev_head + 1) % vfmreg_ctx::MATCH_EVENTS;
    VFM_CUDA(cudaEventRecord(ev, lane));
    VFM_CUDA(cudaStreamWaitEvent(ks, ev, 0));
    ctx->stream = ks;
  }
  group_begin(ctx, grp);
  int rc_launch = VFMREG_OK;
  if (plan.version == 3) {
    if (dp <= A_MAX_KB * TBK)
      match_tc3_kernel<true><<<plan.grid, TC_THREADS, S3_TOTAL, ctx->stream>>>(map_a, map_b, P);
    else
      match_tc3_kernel<false><<<plan.grid, TC_THREADS, S3_TOTAL, ctx->stream>>>(map_a, map_b, P);
    rc_launch = launch_check(ctx, "match_tc3_kernel");
  } else if (plan.version == 2) {
    match_tc2_kernel<<<plan.grid, TC_THREADS, S2_TOTAL, ctx->stream>>>(map_a, map_b, P);
    rc_launch = launch_check(ctx, "match_tc2_kernel");
  } else {
    match_tc_kernel<<<plan.grid, TC_THREADS, SMEM_TOTAL, ctx->stream>>>(map_a, map_b, P);
    rc_launch = launch_check(ctx, "match_tc_kernel");
  }
  group_end(ctx, grp, 1);
  if (ks) {
    ctx->stream = lane;
    cudaEvent_t ev = ctx->match_ev[ctx->match_ev_head];
    ctx->match_ev_head = (ctx->match_ev_head + 1) % vfmreg_ctx::MATCH_EVENTS;
    VFM_CUDA(cudaEventRecord(ev, ks));
    VFM_CUDA(cudaStreamWaitEvent(lane, ev, 0));
  }
  VFM_TRY(rc_launch);
  const long long entries = (long long)n * plan.slots;
  const bool fold_pick = (sec == nullptr);   // only the best match is requested: the pick is folded into the re-score
  rerank_select_kernel<<<ceil_div(entries, 128), 128, 0, ctx->stream>>>((int)n, n_dev, plan.slots, P.top1, nz_a, cand_v, cand_n, slot_top2,
                                                                       work_list, work_count, row_overflow);
  VFM_TRY(launch_check(ctx, "rerank_select_kernel"));
  if (dp <= 512)
    rerank_dot_kernel<8><<<ctx->sm_count * 16, 128, 0, ctx->stream>>>(a32, b32, dp, plan.slots, cand_v, cand_i, work_list, work_count,
                                                                      fold_pick ? row_key : nullptr);
  else
    rerank_dot_kernel<16><<<ctx->sm_count * 16, 128, 0, ctx->stream>>>(a32, b32, dp, plan.slots, cand_v, cand_i, work_list, work_count,
                                                                      fold_pick ? row_key : nullptr);
  VFM_TRY(launch_check(ctx, "rerank_dot_kernel"));
  if (fold_pick) {
    rerank_finish_kernel<<<ceil_div(n, 128), 128, 0, ctx->stream>>>((int)n, n_dev, nz_a, row_key, row_overflow, idx, best, redo_list, redo_count);
    VFM_TRY(launch_check(ctx, "rerank_finish_kernel"));
  } else {
    rerank_pick_kernel<<<ceil_div(n, 128), 128, 0, ctx->stream>>>((int)n, n_dev, (int)m, plan.slots, nz_a, cand_v, cand_i, cand_n, idx, best, sec,
                                                                 redo_list, redo_count);
    VFM_TRY(launch_check(ctx, "rerank_pick_kernel"));
  }
  exact_rows_kernel<<<ctx->sm_count, 256, 0, ctx->stream>>>(a32, b32, (int)m, dp, redo_list, redo_count, idx, best, sec);
  VFM_TRY(launch_check(ctx, "exact_rows_kernel"));
  if (want_dbg) {
    static long long host[2048 * 8];
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpy(host, dbg_dev, sizeof(host), cudaMemcpyDeviceToHost);
    double acc[8] = {0};
    const int mma_ctas = plan.version == 3 ? plan.grid / 2 : plan.grid;   // version 3: only leaders run the MMA loop
    for (int c = 0; c < plan.grid; ++c)
      for (int k = 0; k < 8; ++k) acc[k] += (double)host[c * 8 + k] / (k < 4 ? mma_ctas : plan.grid);
    if (host[1200]) {
      const long long t0 = host[1200];
      for (int i = 0; i < 16; ++i) {
        const long long* r = host + 1200 + i * 16;
        fprintf(stderr, "[tc dbg] tile %2d: mma got tempty %7lld, commit issued %7lld | epi seen/released", i + 8, r[0] - t0, r[1] - t0);
        for (int w = 0; w < 4; ++w) fprintf(stderr, " w%d %7lld/%7lld", w, r[2 + 2 * w] - t0, r[3 + 2 * w] - t0);
        fprintf(stderr, "\n");
      }
    }
#ifdef VFM_SCAN_STATS
    {
      unsigned long long st[4], zero[4] = {0, 0, 0, 0};
      cudaMemcpyFromSymbol(st, g_scan_stats, sizeof(st));
      cudaMemcpyToSymbol(g_scan_stats, zero, sizeof(zero));
      fprintf(stderr, "[tc dbg] scan stats: warp-chunks %llu, with a lane above threshold %llu (%.1f%%), lane-chunks with several hits %llu (%.2f%%), lane-chunks with a hit %llu\n",
              st[0], st[1], 100.0 * st[1] / (st[0] ? st[0] : 1), st[2], 100.0 * st[2] / (st[0] ? st[0] : 1), st[3]);
    }
#endif
    fprintf(stderr, "[tc dbg] v%d grid=%d tiles/cta=%.1f | mma warp: total %.0f cyc, wait tempty %.0f, wait full %.0f | epi warp: total %.0f, wait tfull %.0f | producer: total %.0f, wait empty %.0f\n",
            plan.version, plan.grid, acc[3], acc[0], acc[1], acc[2], acc[4], acc[5], acc[7], acc[6]);
  }
  return VFMREG_OK;
}

}  // namespace vfm
