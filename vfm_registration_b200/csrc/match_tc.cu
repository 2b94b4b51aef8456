// Descriptor top-2 search on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), with results bit-identical to the
// exact fp32 kernel (match_simt.cu) and to oracle/c.
//
// Reference: faiss IndexFlatIP::search(k=1) at VoxelHashMap.cpp:486-495 (+ runner-up for the ratio test).
//
// Two phases
//  1. candidate search -- fp16 x fp16 -> fp32 GEMM of the renormalised descriptors, S~ = A~ B~^T, never written to memory.
//       match_tc3_kernel (default)  CTA pairs, ONE tcgen05.mma.cta_group::2 (M256 N256 K16) per step drives both SMs; each
//                                   CTA holds its 128 query rows (resident for D <= 384) and half of the 256-column tile
//       match_tc_kernel             single CTAs, both operands streamed (fewer than two row blocks)
//     Persistent CTAs walk contiguous spans of (row block, 256-column tile) units in row-major order.  Warp roles: warp 0 =
//     TMA producer (cp.async.bulk.tensor, 128B swizzle, mbarrier ring of 64-wide K chunks), warp 1 = TMEM allocator + MMA
//     issuer (one elected lane, accumulators double-buffered in TMEM: 2 x 256 columns), warps 2.. = epilogue (tcgen05.ld
//     32x32b: one query row per thread).  Each epilogue thread keeps the running approximate best (and runner-up when the
//     caller needs it) of its row and records every column whose approximate score is within `margin` of it in a small
//     shared-memory list (compacted when full); the lists are flushed per (row block, span) "slot".  The row count may
//     live on the device (TcParams.n_dev: the pruned reverse search of register()'s mutual check), rows may start
//     from a known lower bound of their best score (TcParams.seed), and a caller that discards matches below a cosine gate
//     (register() with min_cos) passes that gate as a floor of the recording threshold (TcParams.floor): rows whose best
//     is below the gate then record nothing and report "no match" (index -1, score -inf), which the gate drops anyway.
//  2. re-rank -- rerank_rows_kernel, one warp per query row: rebuild the row's final threshold from the per-slot
//     approximate top-2, drop the candidates below it, recompute the survivors (typically 1-3) in the canonical fp32 order
//     (fmaf chain over k ascending, from the fp32 rows; 16 lanes per candidate hand the accumulator on by shuffle) and pick
//     the exact best / runner-up, lowest index on ties.  A list that overflowed (pathological ties) keeps its largest
//     entries and the largest score it dropped; only if that score could still matter is the row scanned exactly.
//
// Why the result is exact: |S~ - S| <= eps with eps bounded below; the exact best and runner-up of a span both have
// S~ >= (final approx runner-up of that span) - 2 eps, the recording threshold only ever rises, so both are always in
// the list; margin = 2 eps + slack.  eps for unit-norm rows: fp16 input rounding 2^-10 (1 + 2^-12) ||a|| ||b||
// + fp16 subnormal inputs (< 5e-5) + tensor-core fp32 accumulation (< D 2^-23) + canonical fp32 chain (< D 2^-24)
// < 1.2e-3 for D <= 1024; MARGIN = 3e-3.  Only valid for renormalised inputs, which is what the caller guarantees.
// The gate floor is min_cos - MARGIN: every column whose exact score reaches the gate has S~ >= min_cos - eps > floor.
//
// Bound: tensor pipe (2 N M D flop per launch); DESIGN.md 4.1 has the measured fractions.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

namespace vfm {

constexpr int TBM = 128, TBN = 256, TBK = 64, UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int NBUF = 512 / TBN;          // accumulator buffers in TMEM
constexpr int CAP = 16;                 // candidate list entries per (row, slot)
constexpr float MARGIN = 3e-3f;
// Epilogue warps, a template parameter EW of the kernels: 4 (one per TMEM lane quarter) or 8 (warps w and w+4 share a lane
// quarter -- the hardware ties lanes to warp_id % 4 -- and split the tile's columns; each (row, column group) then keeps its own
// candidate list, i.e. HALVES device slots per (row, CTA span)).  Measured (tools/bench_kernels.py modes, 10k x 50k x 384): 8
// warps take top-2 mode from 0.322 to 0.302 ms (0.73 -> 0.78 of the measured bf16 peak); in top-1 mode with the gate floor
// the two are equal, and inside a batch the 4 extra warps take issue slots from the neighbouring lanes' small kernels
// (-3 % pairs/s).  So: top-2 searches run EW = 8, top-1 searches EW = 4.
template <int EW>
struct Epi {
  static constexpr int WARPS = EW, THREADS = EW * 32;
  static constexpr int HALVES = EW / 4;            // column groups per tile
  static constexpr int COLS = TBN / HALVES;        // columns per warp
  static constexpr int TC_THREADS = 64 + THREADS;
  static constexpr uint32_t RING_BYTES = THREADS * CAP * 4;
};
constexpr uint32_t A_STAGE_BYTES = TBM * TBK * 2, B_STAGE_BYTES = TBN * TBK * 2;
// the single-CTA kernel (fewer than two row blocks) always runs 4 epilogue warps
constexpr int EW1 = 4;
constexpr uint32_t SMEM_RING_V = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
constexpr uint32_t SMEM_RING_I = SMEM_RING_V + Epi<EW1>::RING_BYTES;
constexpr uint32_t SMEM_BARS = SMEM_RING_I + Epi<EW1>::RING_BYTES;
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
  float4* slot_top2;     // [n][slots]: approximate (best, runner-up) of the slot's column span, largest score dropped on overflow
  int top1;              // 1: only the best match is needed (runner-up value not requested): threshold = best - margin
  const int* n_dev;      // optional: the number of query rows actually present (<= n), produced earlier on the stream
  const float* seed;     // optional (top1 only): per query row a known lower bound of its exact best score (the pruned
                         // reverse search knows <b_j, a_i> = sim01[i]); recording starts at seed - margin instead of -inf
  float floor;           // top1 only: recording never starts below this (cosine gate - margin), -inf = none
};


// number of query rows of this launch: the host-side bound, or the device-side count when the caller's row list was
// compacted on the GPU (the pruned reverse search of the mutual check)
__device__ __forceinline__ int tc_rows(const TcParams& P) { return P.n_dev ? min(__ldg(P.n_dev), P.n) : P.n; }

// per-thread epilogue state of one query row inside one span
struct EpiRow {
  float best, second, thr;
  float lost;   // largest approximate score that had to be dropped because the list was full (-inf: none)
  int cnt;
};

__device__ __forceinline__ void epi_row_begin(const TcParams& P, EpiRow& s, int row, int n_rows) {
  const bool active = (row < n_rows) && (P.nz[row] != 0);
  s.best = s.second = -INFINITY;
  float t = -INFINITY;
  if (P.top1) {
    if (P.seed && row < n_rows) t = __ldg(P.seed + row) - MARGIN;
    t = fmaxf(t, P.floor);
  }
  s.thr = active ? t : INFINITY;   // inactive rows (padding, all-zero queries) never record
  s.cnt = 0;
  s.lost = -INFINITY;
}

// slot of (row block, span): spans are numbered from the first CTA (cluster) whose span contains the row block's first unit
template <int EW>
__device__ __forceinline__ void flush_slot(const TcParams& P, int n_rows, int row, long long first_unit, long long total_units,
                                           long long n_workers, long long worker, int half, int etid, const float* ring_v,
                                           const int* ring_i, const EpiRow& s) {
  if (row >= n_rows) return;
  const long long g = min(n_workers, total_units);   // workers that own a (non-empty) span
  long long c0 = (first_unit * g) / total_units;
  while ((total_units * (c0 + 1)) / g <= first_unit) ++c0;
  while (c0 > 0 && (total_units * c0) / g > first_unit) --c0;
  const int slot = (int)(worker - c0) * Epi<EW>::HALVES + half;
  const long long o = ((long long)row * P.slots + slot);
  P.cand_n[o] = s.cnt | (s.lost > -INFINITY ? (1 << 30) : 0);
  P.slot_top2[o] = make_float4(s.best, s.second, s.lost, 0.0f);
  for (int e = 0; e < s.cnt; ++e) {
    P.cand_v[o * CAP + e] = ring_v[e * Epi<EW>::THREADS + etid];
    P.cand_i[o * CAP + e] = ring_i[e * Epi<EW>::THREADS + etid];
  }
}

// Record one candidate (approximate score v of column `col`) in the thread's list and raise the recording threshold.
template <bool TOP1, int EW>
__device__ __forceinline__ void push_candidate(float v, int col, int etid, float* ring_v, int* ring_i, EpiRow& s) {
  if (s.cnt == CAP) {  // compact: keep what is still above the (risen) threshold
    int w = 0;
#pragma unroll 1
    for (int e = 0; e < CAP; ++e) {
      const float ev = ring_v[e * Epi<EW>::THREADS + etid];
      const int ei = ring_i[e * Epi<EW>::THREADS + etid];
      if (ev > s.thr) {
        ring_v[w * Epi<EW>::THREADS + etid] = ev;
        ring_i[w * Epi<EW>::THREADS + etid] = ei;
        ++w;
      }
    }
    s.cnt = w;
  }
  if (s.cnt < CAP) {
    ring_v[s.cnt * Epi<EW>::THREADS + etid] = v;
    ring_i[s.cnt * Epi<EW>::THREADS + etid] = col;
    ++s.cnt;
  } else {
    // CAP entries within the margin of the threshold and one more: keep the CAP largest and remember the largest score
    // that was dropped -- the re-rank only needs the exact fallback if that score could still matter at the end
    int lo = 0;
    float lo_v = ring_v[etid];
#pragma unroll 1
    for (int e = 1; e < CAP; ++e) {
      const float ev = ring_v[e * Epi<EW>::THREADS + etid];
      if (ev < lo_v) {
        lo_v = ev;
        lo = e;
      }
    }
    if (v > lo_v) {
      ring_v[lo * Epi<EW>::THREADS + etid] = v;
      ring_i[lo * Epi<EW>::THREADS + etid] = col;
      s.lost = fmaxf(s.lost, lo_v);
    } else {
      s.lost = fmaxf(s.lost, v);
    }
  }
  if (v > s.best) {
    s.second = s.best;
    s.best = v;
  } else if (v > s.second) {
    s.second = v;
  }
  s.thr = fmaxf(s.thr, (TOP1 ? s.best : s.second) - MARGIN);   // never below the gate floor / the seed
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
// Fast path (whole warp): four group maxima, their maximum, one compare against the row's recording threshold and one
// warp vote.  Only when some lane clears its threshold does the warp build the per-lane bit mask of qualifying columns
// (straight-line code, no divergence); lanes with a hit then push -- the usual case, exactly one column, is a single push of
// (chunk maximum, position); several qualifying columns are pushed in ascending column order through a select tree.
template <bool TOP1, int EW>
__device__ __forceinline__ void scan_chunk(const uint32_t* r, int c0, int valid, int col_base, int etid, float* ring_v, int* ring_i,
                                           EpiRow& s) {
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
  if (s.thr == -INFINITY && !partial) {
    // Nothing recorded yet in this span: seed the threshold from this chunk.  TOP1: its maximum.  Otherwise a lower bound
    // of its second largest value: the smaller of two disjoint group maxima.
    const float seed = TOP1 ? mx : fminf(fmaxf(g[0], g[2]), fmaxf(g[1], g[3]));
    s.thr = seed - MARGIN;   // strictly below `seed`, so the values that define it are still recorded below
  }
  if (!__any_sync(0xffffffffu, mx > s.thr)) return;
  uint32_t mask = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) mask |= (__uint_as_float(r[i]) > s.thr) ? (1u << i) : 0u;
  if (partial) mask &= (1u << (valid - c0)) - 1u;
  if (mask == 0) return;   // lanes without a hit wait at the reconvergence point
  if (!partial && (mask & (mask - 1)) == 0) {
    // exactly one qualifying column: it is the chunk maximum
    push_candidate<TOP1, EW>(mx, col_base + c0 + __ffs(mask) - 1, etid, ring_v, ring_i, s);
  } else {
#pragma unroll 1
    while (mask) {
      const int i = __ffs(mask) - 1;
      mask &= mask - 1;
      const float v = pick32(r, i);
      if (v > s.thr) push_candidate<TOP1, EW>(v, col_base + c0 + i, etid, ring_v, ring_i, s);
    }
  }
}

// The epilogue of one warp's share of an accumulator tile (Epi<EW>::COLS columns starting at TMEM address t_addr /
// database column col_base): 32-column chunks, the next one in flight while the current one is scanned.
template <bool TOP1, int EW>
__device__ __forceinline__ void epilogue_half(uint32_t t_addr, int col_base, int m, int etid, float* ring_v, int* ring_i, EpiRow& s) {
  const int valid = min(Epi<EW>::COLS, m - col_base);   // may be <= 0 for the padded part of the last tile
  uint32_t ra[32], rbuf[32];
  tc_ld32(t_addr, ra);
#pragma unroll 1
  for (int c = 0; c < Epi<EW>::COLS / 32; c += 2) {
    tc_wait_ld();
    tc_ld32(t_addr + (c + 1) * 32, rbuf);  // in flight while chunk c is scanned
    scan_chunk<TOP1, EW>(ra, c * 32, valid, col_base, etid, ring_v, ring_i, s);
    __syncwarp();
    tc_wait_ld();
    if (c + 2 < Epi<EW>::COLS / 32) tc_ld32(t_addr + (c + 2) * 32, ra);
    scan_chunk<TOP1, EW>(rbuf, (c + 1) * 32, valid, col_base, etid, ring_v, ring_i, s);
    __syncwarp();
  }
}

template <int EW>
__device__ __forceinline__ void epilogue_tile(bool top1, uint32_t t_addr, int col_base, int m, int etid, float* ring_v, int* ring_i,
                                              EpiRow& s) {
  if (top1)
    epilogue_half<true, EW>(t_addr, col_base, m, etid, ring_v, ring_i, s);
  else
    epilogue_half<false, EW>(t_addr, col_base, m, etid, ring_v, ring_i, s);
}

// ---------------------------------------------------------------------------------------------------------------------
// Single-CTA kernel: both operands streamed.  Serves searches with fewer than two 128-row query blocks.
__global__ void __launch_bounds__(Epi<EW1>::TC_THREADS, 1)
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
      mbar_init(tempty0 + 8 * b, Epi<EW1>::WARPS);
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
    uint32_t stage = 0, phase = 0;
    for (long long t = t_begin; t < t_end; ++t) {
      const int rb = (int)(t / P.col_tiles), ct = (int)(t % P.col_tiles);
#pragma unroll 1
      for (int kb = 0; kb < P.kb; ++kb) {
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(full0 + 8 * stage, A_STAGE_BYTES + B_STAGE_BYTES);
          tma_load_2d(sA + stage * A_STAGE_BYTES, &map_a, full0 + 8 * stage, kb * TBK, rb * TBM);
          tma_load_2d(sB + stage * B_STAGE_BYTES, &map_b, full0 + 8 * stage, kb * TBK, ct * TBN);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp-uniform loop, one elected lane issues) =====
    uint32_t stage = 0, phase = 0;
    long long it = 0;
    const uint64_t da0 = umma_desc_k_sw128(sA), db0 = umma_desc_k_sw128(sB);
    for (long long t = t_begin; t < t_end; ++t, ++it) {
      const uint32_t buf = (uint32_t)(it % NBUF);
      mbar_wait(tempty0 + 8 * buf, (uint32_t)((it / NBUF) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * TBN;
#pragma unroll 1
      for (int kb = 0; kb < P.kb; ++kb) {
        mbar_wait(full0 + 8 * stage, phase);
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
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue: TMEM lane quarter = warp % 4, column group = (warp - 2) / 4 =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int tid = q * 32 + lane;        // row inside the block == TMEM lane
    const int etid = half * 128 + tid;    // index among the epilogue threads
    int cur_rb = -1;
    EpiRow s;
    s.best = s.second = -INFINITY;
    s.thr = INFINITY;
    s.cnt = 0;
    s.lost = -INFINITY;
    long long it = 0;
    for (long long t = t_begin; t < t_end; ++t, ++it) {
      const int rb = (int)(t / P.col_tiles), ct = (int)(t % P.col_tiles);
      if (rb != cur_rb) {
        if (cur_rb >= 0)
          flush_slot<EW1>(P, n_rows, cur_rb * TBM + tid, (long long)cur_rb * P.col_tiles, total_tiles, gridDim.x, blockIdx.x, half, etid,
                     ring_v, ring_i, s);
        cur_rb = rb;
        epi_row_begin(P, s, rb * TBM + tid, n_rows);
      }
      const uint32_t buf = (uint32_t)(it % NBUF);
      mbar_wait(tfull0 + 8 * buf, (uint32_t)((it / NBUF) & 1));
      tc_fence_after();
      epilogue_tile<EW1>(P.top1 != 0, tmem_base + ((uint32_t)(q * 32) << 16) + buf * TBN + half * Epi<EW1>::COLS, ct * TBN + half * Epi<EW1>::COLS,
                    P.m, etid, ring_v, ring_i, s);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
    }
    if (cur_rb >= 0)
      flush_slot<EW1>(P, n_rows, cur_rb * TBM + tid, (long long)cur_rb * P.col_tiles, total_tiles, gridDim.x, blockIdx.x, half, etid, ring_v,
                 ring_i, s);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair MMA (tcgen05.mma.cta_group::2).  One instruction issued by the leader CTA computes a 256 x 256 x 16
// block on the tensor cores of both SMs of the pair: each CTA holds its own 128 query rows (A) and HALF of the 256-column
// database tile (B) in its own shared memory, and receives the 128 x 256 accumulator rows of its query block in its own
// TMEM.  Against two independent M128 N256 MMAs the shared-memory operand traffic per MMA drops from 12 KB to 8 KB per SM
// -- the measured limiter of the single-CTA kernel (profiles/r1_kernel_bench.txt) -- and every database byte still crosses
// L2 -> SM once per pair without multicast.
//   RESIDENT (dp <= 384): the query block stays in shared memory for the whole row of column tiles; only B is streamed.
//   otherwise (dp up to 1024): A and B chunks are streamed together.
// Barriers: `full` / `afull` / `tempty` live in the leader (both CTAs' TMA loads credit the leader's barriers through
// .cta_group::2 loads; the peer's epilogue warps arrive remotely); `empty` / `aempty` / `tfull` exist in both CTAs and are
// signalled by multicast tcgen05.commit.
constexpr int A_MAX_KB = 6;  // resident query block: up to 384 columns
constexpr uint32_t B_HALF_BYTES = B_STAGE_BYTES / 2;
constexpr int STAGES3 = 6;
constexpr uint32_t S3_A = 0;                                   // RESIDENT: dp/64 chunks; streaming: STAGES3 chunks
constexpr uint32_t S3_B = A_MAX_KB * A_STAGE_BYTES;            // STAGES3 x (128 database rows x 64 k) = 16 KB each
constexpr uint32_t S3_RING_V = S3_B + STAGES3 * B_HALF_BYTES;
template <int EW> constexpr uint32_t s3_ring_i() { return S3_RING_V + Epi<EW>::RING_BYTES; }
template <int EW> constexpr uint32_t s3_bars() { return s3_ring_i<EW>() + Epi<EW>::RING_BYTES; }
template <int EW> constexpr uint32_t s3_total() { return s3_bars<EW>() + 256 + 1024; }
constexpr uint32_t IDESC3 = umma_idesc_f16(2 * TBM, TBN, 0);
static_assert(TBN == 256 && NBUF == 2, "the CTA-pair kernel is written for 256-column tiles");
static_assert(STAGES3 <= A_MAX_KB, "streamed A chunks reuse the resident region");
static_assert(s3_total<8>() <= 232448, "shared memory budget of one SM");

template <bool RESIDENT, int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Epi<EW>::TC_THREADS, 1)
    match_tc3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base + S3_A, sB = base + S3_B;
  float* ring_v = reinterpret_cast<float*>(smem + S3_RING_V);
  int* ring_i = reinterpret_cast<int*>(smem + s3_ring_i<EW>());
  const uint32_t bars = base + s3_bars<EW>();
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES3, tfull0 = bars + 16 * STAGES3, tempty0 = tfull0 + 8 * NBUF;
  const uint32_t afull = tempty0 + 8 * NBUF, aempty = afull + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + s3_bars<EW>() + 16 * STAGES3 + 16 * NBUF + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_cta_rank();
  const long long clusters = gridDim.x >> 1, cid = blockIdx.x >> 1;
  TraceScope trace(P.n_dev ? 21 : 20);   // trace build only: 20 = full search, 21 = pruned reverse search
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
    for (int s = 0; s < STAGES3; ++s) {
      mbar_init(full0 + 8 * s, 1);    // leader: one arrive.expect_tx per phase, bytes from both CTAs' loads
      mbar_init(empty0 + 8 * s, 1);   // both CTAs: multicast commit from the leader's MMA warp
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(tfull0 + 8 * b, 1);                // both CTAs: multicast commit
      mbar_init(tempty0 + 8 * b, 2 * Epi<EW>::WARPS);   // leader: epilogue warps of both CTAs
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
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
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
  } else if (warp == 1) {
    // ===== MMA issuer: the leader CTA only =====
    if (rank == 0) {
      uint32_t stage = 0, phase = 0, aphase = 0;
      long long it = 0;
      int cur_rp = -1;
      const uint64_t da0 = umma_desc_k_sw128(sA), db0 = umma_desc_k_sw128(sB);
      for (long long u = u_begin; u < u_end; ++u, ++it) {
        const int rp = (int)(u / P.col_tiles);
        const uint32_t buf = (uint32_t)(it % NBUF);
        mbar_wait(tempty0 + 8 * buf, (uint32_t)((it / NBUF) & 1) ^ 1);
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
          mbar_wait(full0 + 8 * stage, phase);
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
    }
  } else {
    // ===== epilogue: this CTA's 128 rows x 256 columns =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int tid = q * 32 + lane;
    const int etid = half * 128 + tid;
    int cur_rb = -1;
    EpiRow s;
    s.best = s.second = -INFINITY;
    s.thr = INFINITY;
    s.cnt = 0;
    s.lost = -INFINITY;
    long long it = 0;
    const uint32_t tempty_leader0 = mapa_cluster(tempty0, 0);
    for (long long u = u_begin; u < u_end; ++u, ++it) {
      const int rp = (int)(u / P.col_tiles), ct = (int)(u % P.col_tiles);
      const int rb = 2 * rp + (int)rank;
      if (rb != cur_rb) {
        if (cur_rb >= 0)
          flush_slot<EW>(P, n_rows, cur_rb * TBM + tid, (long long)(cur_rb >> 1) * P.col_tiles, total_units, clusters, cid, half, etid,
                     ring_v, ring_i, s);
        cur_rb = rb;
        epi_row_begin(P, s, rb * TBM + tid, n_rows);
      }
      const uint32_t buf = (uint32_t)(it % NBUF);
      mbar_wait(tfull0 + 8 * buf, (uint32_t)((it / NBUF) & 1));
      tc_fence_after();
      epilogue_tile<EW>(P.top1 != 0, tmem_base + ((uint32_t)(q * 32) << 16) + buf * TBN + half * Epi<EW>::COLS, ct * TBN + half * Epi<EW>::COLS,
                    P.m, etid, ring_v, ring_i, s);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader0 + 8 * buf);
    }
    if (cur_rb >= 0)
      flush_slot<EW>(P, n_rows, cur_rb * TBM + tid, (long long)(cur_rb >> 1) * P.col_tiles, total_units, clusters, cid, half, etid, ring_v,
                 ring_i, s);
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

// multiset top-2 merge of (x1 >= x2) into (t1 >= t2)
__device__ __forceinline__ void top2_merge(float& t1, float& t2, float x1, float x2) {
  const float lo = fminf(t1, x1);
  t1 = fmaxf(t1, x1);
  t2 = fmaxf(fmaxf(t2, x2), lo);
}

// Exact scan of all columns by one warp, for the (rare) query rows whose candidate list overflowed in a way that matters:
// lane l takes columns 4 (32 t + l) .. + 3 -- four independent canonical fmaf chains per lane -- keeps its exact best (lowest
// index on ties: columns ascend inside a lane and '>' is strict) and runner-up, then the lanes are merged by (score, index).
__device__ __noinline__ void exact_row_scan(const float* __restrict__ ar, const float* __restrict__ b, int m, int dp, float& b1,
                                            int& bi, float& b2) {
  const int lane = threadIdx.x & 31;
  b1 = b2 = -INFINITY;
  bi = 0x7fffffff;
  const float4* a4 = reinterpret_cast<const float4*>(ar);
  for (int j0 = 4 * lane; j0 < m; j0 += 128) {
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const float4* r4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) r4[u] = reinterpret_cast<const float4*>(b + (long long)min(j0 + u, m - 1) * dp);
    for (int k = 0; k < dp / 4; ++k) {
      const float4 x = __ldg(a4 + k);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 y = __ldg(r4[u] + k);
        acc[u] = fmaf(x.x, y.x, acc[u]);
        acc[u] = fmaf(x.y, y.y, acc[u]);
        acc[u] = fmaf(x.z, y.z, acc[u]);
        acc[u] = fmaf(x.w, y.w, acc[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (j0 + u >= m) continue;
      if (acc[u] > b1) {
        b2 = b1;
        b1 = acc[u];
        bi = j0 + u;
      } else if (acc[u] > b2) {
        b2 = acc[u];
      }
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const float y1 = __shfl_xor_sync(0xffffffffu, b1, off), y2 = __shfl_xor_sync(0xffffffffu, b2, off);
    const int yi = __shfl_xor_sync(0xffffffffu, bi, off);
    const bool y_wins = (y1 > b1) || (y1 == b1 && yi < bi);
    const float lo = y_wins ? b1 : y1;
    b2 = fmaxf(fmaxf(b2, y2), lo);
    if (y_wins) {
      b1 = y1;
      bi = yi;
    }
  }
}

// Re-rank: one warp per query row.
//   pass 1  the row's final recording threshold from the per-slot approximate top-2 (second largest of all slot bests /
//           runner-ups, or the largest in top-1 mode, minus the margin);
//   pass 2  the row's candidate entries (slots x CAP, 32 at a time) are filtered against it; the survivors are re-scored two
//           at a time, 16 lanes per candidate: lane l owns the l-th contiguous segment of the two rows (dp / 16 elements, at
//           most 4 PER_MAX floats), so every load of a candidate is in flight at once; the canonical fmaf chain over k ascending
//           is then walked lane by lane, the accumulator handed on by shuffle -- the same operation order as one thread running
//           the whole chain;
//   pick    exact best (lowest index on ties) and runner-up value (multiset) -> idx / best / sec.
// Rows whose list overflowed in a way that matters are scanned exactly by their warp (exact_row_scan).  With a gate floor (`floor_mode`) a row may have no candidate at all:
// its best is below the caller's gate and it reports index -1 / score -inf.
template <int PER_MAX>
__global__ void __launch_bounds__(128)
    rerank_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, int dp, int n, const int* __restrict__ n_dev, int m,
                       int slots, int top1, int floor_mode, const uint8_t* __restrict__ nz, const float* __restrict__ cand_v,
                       const int* __restrict__ cand_i, const int* __restrict__ cand_n, const float4* __restrict__ slot_top2,
                       int32_t* __restrict__ idx, float* __restrict__ best, float* __restrict__ sec) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n_dev) n = min(n, __ldg(n_dev));
  if (row >= n) return;   // whole warp
  if (!nz[row]) {  // all-zero query: every canonical inner product is exactly +0 -> lowest index wins
    if (lane == 0) {
      idx[row] = 0;
      if (best) best[row] = 0.0f;
      if (sec) sec[row] = (m > 1) ? 0.0f : -INFINITY;
    }
    return;
  }
  const long long srow = (long long)row * slots;
  // ---- one round trip: per-slot counts and approximate top-2, and (speculatively) the first 64 candidate entries
  const int entries = slots * CAP;
  int cn_l = 0;
  float4 tt_l = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.0f);
  if (lane < slots) {
    cn_l = cand_n[srow + lane];
    tt_l = slot_top2[srow + lane];
  }
  float pv0 = 0.0f, pv1 = 0.0f;
  int pi0 = 0, pi1 = 0;
  if (lane < entries) {
    pv0 = cand_v[srow * CAP + lane];
    pi0 = cand_i[srow * CAP + lane];
  }
  if (32 + lane < entries) {
    pv1 = cand_v[srow * CAP + 32 + lane];
    pi1 = cand_i[srow * CAP + 32 + lane];
  }
  // ---- pass 1
  float t1 = -INFINITY, t2 = -INFINITY, lost = -INFINITY;
  for (int s = lane; s < slots; s += 32) {
    const int c = (s == lane) ? cn_l : cand_n[srow + s];
    if ((c & 0xFFFF) == 0 && !((c >> 30) & 1)) continue;
    const float4 t = (s == lane) ? tt_l : slot_top2[srow + s];
    top2_merge(t1, t2, t.x, t.y);
    if ((c >> 30) & 1) lost = fmaxf(lost, t.z);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const float o1 = __shfl_xor_sync(FULL, t1, off), o2 = __shfl_xor_sync(FULL, t2, off);
    top2_merge(t1, t2, o1, o2);
    lost = fmaxf(lost, __shfl_xor_sync(FULL, lost, off));
  }
  const float thr = (top1 ? t1 : t2) - MARGIN;
  // a list that overflowed dropped its smallest entries; that only matters if one of them reaches the final threshold
  // (the same test that discards recorded candidates below)
  if (lost > -INFINITY && lost >= thr) {
    float e1, e2;
    int ei;
    exact_row_scan(a + (long long)row * dp, b, m, dp, e1, ei, e2);
    if (lane == 0) {
      idx[row] = (ei == 0x7fffffff) ? 0 : ei;
      if (best) best[row] = e1;
      if (sec) sec[row] = e2;
    }
    return;
  }
  // ---- pass 2
  constexpr int L = 16;
  const int gl = lane & (L - 1), hb = lane & ~(L - 1), h = lane >> 4;
  const int per = dp / (4 * L);   // float4 per lane and row (dp % 64 == 0, dp <= 64 PER_MAX)
  const float4* a4 = reinterpret_cast<const float4*>(a + (long long)row * dp) + gl * per;
  float4 av[PER_MAX];
  bool a_loaded = false;
  float b1 = -INFINITY, b2 = -INFINITY;   // per half-warp: exact best / runner-up over the candidates it scored
  int bi = 0x7fffffff;
  for (int e0 = 0; e0 < entries; e0 += 32) {
    const int e = e0 + lane;
    bool keep = false;
    int col = 0;
    if (e0 < 64) {   // prefetched above; the slot of entry e is e / CAP < 4 <= 32 lanes
      const int cnt = __shfl_sync(FULL, cn_l, (e / CAP) & 31) & 0xFFFF;
      const float v = e0 ? pv1 : pv0;
      if (e < entries && (e % CAP) < cnt && v >= thr) {
        keep = true;
        col = e0 ? pi1 : pi0;
      }
    } else if (e < entries) {
      const int cnt = cand_n[srow + e / CAP] & 0xFFFF;
      if ((e % CAP) < cnt && cand_v[srow * CAP + e] >= thr) {
        keep = true;
        col = cand_i[srow * CAP + e];
      }
    }
    unsigned live = __ballot_sync(FULL, keep);
    if (live && !a_loaded) {   // warp-uniform
#pragma unroll
      for (int u = 0; u < PER_MAX; ++u)
        if (u < per) av[u] = __ldg(a4 + u);
      a_loaded = true;
    }
    while (live) {   // warp-uniform: two survivors per round, one per half-warp
      const int s0 = __ffs(live) - 1;
      live &= live - 1;
      int s1 = -1;
      if (live) {
        s1 = __ffs(live) - 1;
        live &= live - 1;
      }
      const int src = h ? s1 : s0;
      const bool mine = src >= 0;
      const int c = __shfl_sync(FULL, col, mine ? src : 0);
      float4 bv[PER_MAX];
      if (mine) {
        const float4* b4 = reinterpret_cast<const float4*>(b + (long long)c * dp) + gl * per;
#pragma unroll
        for (int u = 0; u < PER_MAX; ++u)
          if (u < per) bv[u] = __ldg(b4 + u);
      }
      float acc = 0.0f;
#pragma unroll 1
      for (int sl = 0; sl < L; ++sl) {
        if (mine && gl == sl) {
#pragma unroll
          for (int u = 0; u < PER_MAX; ++u)
            if (u < per) {
              acc = fmaf(av[u].x, bv[u].x, acc);
              acc = fmaf(av[u].y, bv[u].y, acc);
              acc = fmaf(av[u].z, bv[u].z, acc);
              acc = fmaf(av[u].w, bv[u].w, acc);
            }
        }
        acc = __shfl_sync(FULL, acc, hb + sl);   // lane sl's running sum -> every lane of the half-warp
      }
      if (mine) {
        if (acc > b1 || (acc == b1 && c < bi)) {
          b2 = b1;
          b1 = acc;
          bi = c;
        } else if (acc > b2) {
          b2 = acc;
        }
      }
    }
  }
  // ---- pick: merge the two half-warps
  {
    const float y1 = __shfl_xor_sync(FULL, b1, 16), y2 = __shfl_xor_sync(FULL, b2, 16);
    const int yi = __shfl_xor_sync(FULL, bi, 16);
    const bool y_wins = (y1 > b1) || (y1 == b1 && yi < bi);
    const float lo = y_wins ? b1 : y1;
    b2 = fmaxf(fmaxf(b2, y2), lo);
    if (y_wins) {
      b1 = y1;
      bi = yi;
    }
  }
  if (bi == 0x7fffffff && !floor_mode) {   // cannot happen for a non-zero row; exact scan as a safety net
    exact_row_scan(a + (long long)row * dp, b, m, dp, b1, bi, b2);
    if (bi == 0x7fffffff) bi = 0;
  }
  if (lane == 0) {
    if (bi == 0x7fffffff) {   // floor mode: nothing reaches the caller's gate
      idx[row] = -1;
      if (best) best[row] = -INFINITY;
      if (sec) sec[row] = -INFINITY;
    } else {
      idx[row] = bi;
      if (best) best[row] = b1;
      if (sec) sec[row] = b2;
    }
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
  bool paired;   // CTA pairs (cta_group::2 MMA); single CTAs serve searches with fewer than two row blocks
};

// `dynamic`: n is only an upper bound, the kernels read the row count from device memory.  The grid is sized for n; the
// slot bound must then hold for every smaller total: spans of at least one unit cut a row block's col_tiles consecutive
// units into at most col_tiles + 1 pieces (and spans of at most one unit into at most col_tiles).
static TcPlan tc_plan(vfmreg_ctx* ctx, int64_t n, int64_t m, bool dynamic, int halves) {
  TcPlan p;
  p.row_blocks = ceil_div(n, TBM);
  p.col_tiles = ceil_div(m, TBN);
  p.paired = p.row_blocks >= 2;
  if (p.paired) {
    const int row_pairs = ceil_div(n, 2 * TBM);
    p.total = (long long)row_pairs * p.col_tiles;          // units of (row-block pair, column tile)
    const int clusters = (int)((p.total < ctx->sm_count / 2) ? p.total : ctx->sm_count / 2);
    p.grid = 2 * clusters;
    const long long min_span = p.total / clusters;
    p.slots = dynamic ? p.col_tiles + 1 : (int)((p.col_tiles + min_span - 1) / min_span) + 1;
    if (p.slots > clusters) p.slots = clusters;
    p.slots *= halves;   // column groups per span
    return p;
  }
  p.total = (long long)p.row_blocks * p.col_tiles;
  p.grid = (int)((p.total < ctx->sm_count) ? p.total : ctx->sm_count);
  // a row block of col_tiles consecutive tiles is cut by at most ceil(col_tiles / floor(total/grid)) + 1 spans
  const long long min_span = p.total / p.grid;
  p.slots = dynamic ? p.col_tiles + 1 : (int)((p.col_tiles + min_span - 1) / min_span) + 1;
  if (p.slots > p.grid) p.slots = p.grid;
  p.slots *= 1;        // the single-CTA kernel runs 4 epilogue warps
  return p;
}

size_t match_tc_scratch(vfmreg_ctx* ctx, int64_t n, int64_t m, bool dynamic) {
  const TcPlan p = tc_plan(ctx, n, m, dynamic, 2);   // sized for the 8-warp variant (two column groups per span)
  return 2 * arena_bytes((size_t)n * p.slots * CAP, 4) + arena_bytes((size_t)n * p.slots, 4) +
         arena_bytes((size_t)n * p.slots, 16) + 2048;
}

// a32/b32: renormalised fp32 rows (n x dp), a16/b16: their fp16 copies, nz_a: non-zero flags of the query rows.
// n_dev (optional, device): the number of query rows actually present; n is then the capacity of a16 / a32 / nz_a / idx.
// floor (top-1 mode only, NAN = none): the caller drops matches whose score is below this gate, so rows that cannot reach
// it may report "no match" (index -1, score -inf) -- see the header of this file.
//
// match_tc_begin enqueues the candidate search, match_tc_finish the re-rank.  In batch mode (ctx->match_stream set) the search
// kernel goes to the batch's high-priority search stream -- the lane stream hands over with an event -- and `pending->done`
// is recorded behind it; match_tc_finish makes the lane wait for `wait_for` (the batch passes the event behind the LAST
// search of a group, so that no small kernel runs beside a search) or for `pending->done`.
int match_tc_begin(vfmreg_ctx* ctx, const float* a32, const void* a16, const uint8_t* nz_a, int64_t n, const float* b32,
                   const void* b16, int64_t m, int dp, int32_t* idx, float* best, float* sec, const int* n_dev, const float* seed,
                   float floor, TcPending* pending) {
  VFM_CHECK_ARG(dp % TBK == 0 && dp <= 1024, "match_tc: padded dim %d must be a multiple of %d and <= 1024 (error bound)", dp, TBK);
  VFM_CHECK_ARG(n > 0 && m > 0 && n < (1LL << 30) && m < (1LL << 30), "match_tc: bad sizes");
  const bool top1 = sec == nullptr;
  const int ew = top1 ? 4 : 8;   // epilogue warps of the pair kernel (see Epi)
  const TcPlan plan = tc_plan(ctx, n, m, n_dev != nullptr, ew / 4);
  float* cand_v = arena_take<float>(ctx, (size_t)n * plan.slots * CAP);
  int* cand_i = arena_take<int>(ctx, (size_t)n * plan.slots * CAP);
  float4* slot_top2 = arena_take<float4>(ctx, (size_t)n * plan.slots);
  // slot counts start at zero: a (row, slot) that no span covers stays empty
  const size_t z_bytes = arena_bytes((size_t)n * plan.slots, 4);
  int* cand_n = arena_take<int>(ctx, (size_t)n * plan.slots);
  if (!cand_v || !cand_i || !slot_top2 || !cand_n) {
    set_error("match_tc: scratch arena too small");
    return VFMREG_ERR_ALLOC;
  }
  VFM_CUDA(cudaMemsetAsync(cand_n, 0, z_bytes, ctx->stream));
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
  P.top1 = top1 ? 1 : 0;
  P.n_dev = n_dev;
  P.seed = seed;
  const bool floor_mode = P.top1 && !(floor != floor);
  P.floor = floor_mode ? floor - MARGIN : -INFINITY;
  if (!ctx->tc_attr_set) {   // function attributes are per device; one context per device
    VFM_CUDA(cudaFuncSetAttribute(match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL));
    VFM_CUDA(cudaFuncSetAttribute(match_tc3_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s3_total<4>()));
    VFM_CUDA(cudaFuncSetAttribute(match_tc3_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s3_total<4>()));
    VFM_CUDA(cudaFuncSetAttribute(match_tc3_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s3_total<8>()));
    VFM_CUDA(cudaFuncSetAttribute(match_tc3_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s3_total<8>()));
    ctx->tc_attr_set = true;
  }
  const int grp = n_dev ? GROUP_MATCH_PRUNED : GROUP_MATCH;
  cudaStream_t lane = ctx->stream;
  cudaStream_t ks = ctx->match_stream;
  pending->done = nullptr;
  if (ks) {
    cudaEvent_t ev = ctx->match_ev[ctx->match_ev_head];
    ctx->match_ev_head = (ctx->match_ev_head + 1) % vfmreg_ctx::MATCH_EVENTS;
    VFM_CUDA(cudaEventRecord(ev, lane));
    VFM_CUDA(cudaStreamWaitEvent(ks, ev, 0));
    ctx->stream = ks;
  }
  nvtx_push("match_tc");
  group_begin(ctx, grp);
  int rc_launch = VFMREG_OK;
  if (plan.paired) {
    const bool resident = dp <= A_MAX_KB * TBK;
    if (ew == 4) {
      if (resident)
        match_tc3_kernel<true, 4><<<plan.grid, Epi<4>::TC_THREADS, s3_total<4>(), ctx->stream>>>(map_a, map_b, P);
      else
        match_tc3_kernel<false, 4><<<plan.grid, Epi<4>::TC_THREADS, s3_total<4>(), ctx->stream>>>(map_a, map_b, P);
    } else {
      if (resident)
        match_tc3_kernel<true, 8><<<plan.grid, Epi<8>::TC_THREADS, s3_total<8>(), ctx->stream>>>(map_a, map_b, P);
      else
        match_tc3_kernel<false, 8><<<plan.grid, Epi<8>::TC_THREADS, s3_total<8>(), ctx->stream>>>(map_a, map_b, P);
    }
    rc_launch = launch_check(ctx, "match_tc3_kernel");
  } else {
    match_tc_kernel<<<plan.grid, Epi<EW1>::TC_THREADS, SMEM_TOTAL, ctx->stream>>>(map_a, map_b, P);
    rc_launch = launch_check(ctx, "match_tc_kernel");
  }
  group_end(ctx, grp, 1);
  nvtx_pop();
  if (ks) {
    ctx->stream = lane;
    cudaEvent_t ev = ctx->match_ev[ctx->match_ev_head];
    ctx->match_ev_head = (ctx->match_ev_head + 1) % vfmreg_ctx::MATCH_EVENTS;
    VFM_CUDA(cudaEventRecord(ev, ks));
    pending->done = ev;
  }
  VFM_TRY(rc_launch);
  pending->a32 = a32;
  pending->b32 = b32;
  pending->nz_a = nz_a;
  pending->n = n;
  pending->m = m;
  pending->dp = dp;
  pending->slots = plan.slots;
  pending->top1 = P.top1;
  pending->floor_mode = floor_mode ? 1 : 0;
  pending->n_dev = n_dev;
  pending->cand_v = cand_v;
  pending->cand_i = cand_i;
  pending->cand_n = cand_n;
  pending->slot_top2 = slot_top2;
  pending->idx = idx;
  pending->best = best;
  pending->sec = sec;
  return VFMREG_OK;
}

int match_tc_finish(vfmreg_ctx* ctx, const TcPending& t, cudaEvent_t wait_for) {
  if (wait_for)
    VFM_CUDA(cudaStreamWaitEvent(ctx->stream, wait_for, 0));
  else if (t.done)
    VFM_CUDA(cudaStreamWaitEvent(ctx->stream, t.done, 0));
  GroupScope g_rerank(ctx, GROUP_RERANK, 1);
  const int rows_per_cta = 4;
  const int n = (int)t.n, dp = t.dp;
#define VFM_RERANK(PER)                                                                                                      \
  rerank_rows_kernel<PER><<<ceil_div(n, rows_per_cta), rows_per_cta * 32, 0, ctx->stream>>>(                                 \
      t.a32, t.b32, dp, n, t.n_dev, (int)t.m, t.slots, t.top1, t.floor_mode, t.nz_a, t.cand_v, t.cand_i, t.cand_n,             \
      static_cast<const float4*>(t.slot_top2), t.idx, t.best, t.sec)
  if (dp <= 384)
    VFM_RERANK(6);
  else if (dp <= 512)
    VFM_RERANK(8);
  else if (dp <= 768)
    VFM_RERANK(12);
  else
    VFM_RERANK(16);
#undef VFM_RERANK
  return launch_check(ctx, "rerank_rows_kernel");
}

int match_tc(vfmreg_ctx* ctx, const float* a32, const void* a16, const uint8_t* nz_a, int64_t n, const float* b32,
             const void* b16, int64_t m, int dp, int32_t* idx, float* best, float* sec, const int* n_dev, const float* seed,
             float floor) {
  TcPending t;
  VFM_TRY(match_tc_begin(ctx, a32, a16, nz_a, n, b32, b16, m, dp, idx, best, sec, n_dev, seed, floor, &t));
  return match_tc_finish(ctx, t, nullptr);
}

}  // namespace vfm

VFM_TRACE_ATTACH(vfmreg_trace_attach_match)
