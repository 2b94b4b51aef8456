// tcgen05 / TMEM / TMA / mbarrier PTX wrappers shared by the tensor-core kernels (sm_100a).
// Syntax follows the PTX ISA as used in the CUTLASS sm100 headers (cute/arch/copy_sm100.hpp, mma_sm100_umma.hpp,
// tmem_allocator_sm100.hpp); nothing here is library code.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace vfm {

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile in shared memory, rows of 64 fp16 = 128 B, 128-byte swizzle (what TMA SWIZZLE_128B writes):
// start address >> 4, LBO unused for swizzled K-major, SBO = 8 rows x 128 B = 1024 B, descriptor version 1, layout 2.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;  // LBO = 1 (canonical value for swizzled K-major, ignored by the hardware)
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}


// one lane of a fully converged warp (keeps the surrounding control flow and operands warp-uniform, so descriptors and
// barrier addresses stay in uniform registers instead of being re-broadcast around every tcgen05 / TMA instruction)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// --- programmatic dependent launch (kernels launched with cudaLaunchAttributeProgrammaticStreamSerialization) ---
// launch_dependents: the next kernel on the stream may be scheduled once every CTA of this grid has executed it (or exited);
// wait: blocks until the grids this one depends on have completed and their memory is visible.  Code before the wait must
// not touch global memory that another kernel writes or reads.  Both are no-ops in a kernel launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// L2 prefetch of this CTA's slice (part `idx` of `parts`) of a byte range another kernel will stream soon (the weights of the
// next GEMM: at a few images per batch every GEMM would otherwise start by pulling its weights out of HBM); one thread
__device__ __forceinline__ void l2_prefetch_slice(const char* ptr, unsigned long long bytes, unsigned idx, unsigned parts) {
  if (!ptr || bytes == 0) return;
  const unsigned long long chunk = 16384ull;
  const unsigned long long n_chunks = (bytes + chunk - 1) / chunk;
  for (unsigned long long c = idx; c < n_chunks; c += parts) {
    const unsigned long long off = c * chunk;
    const unsigned sz = (unsigned)((bytes - off < chunk ? bytes - off : chunk) & ~15ull);
    if (sz) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr + off), "r"(sz) : "memory");
  }
}

// --- in-situ timeline (tools/vit_trace.py; built only with -DVFM_TRACE into libvfmreg_b200_trace.so) ---
// Thread 0 of every CTA appends (kind | block << 8 | smid << 40, entry, after griddepcontrol.wait, exit) in globaltimer
// nanoseconds: the timeline of a CUDA-graph replay with programmatic dependent launches, which no profiler here shows
// (ncu serialises the launches and flushes the caches between them).  Compiles to nothing in the product build.
#ifdef VFM_TRACE
struct TraceBuf {
  unsigned long long* rec;
  unsigned int* cursor;
  unsigned int cap;
};
static __device__ TraceBuf g_trace;
__device__ __forceinline__ unsigned long long trace_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
struct TraceScope {
  unsigned long long t0, t1;
  int kind;
  __device__ __forceinline__ explicit TraceScope(int k) : kind(k) { t0 = t1 = trace_now(); }
  __device__ __forceinline__ void waited() { t1 = trace_now(); }
  __device__ __forceinline__ void end() {
    if (threadIdx.x != 0 || !g_trace.rec) return;
    const unsigned int i = atomicAdd(g_trace.cursor, 1u);
    if (i >= g_trace.cap) return;
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const unsigned long long blk = blockIdx.x + (unsigned long long)gridDim.x * (blockIdx.y + (unsigned long long)gridDim.y * blockIdx.z);
    g_trace.rec[4 * i + 0] = (unsigned long long)kind | (blk << 8) | ((unsigned long long)smid << 40);
    g_trace.rec[4 * i + 1] = t0;
    g_trace.rec[4 * i + 2] = t1;
    g_trace.rec[4 * i + 3] = trace_now();
  }
};
// timestamps of one thread of CTA 0 (kinds >= 100), kept in local memory while the kernel runs (a store does not stall the
// thread) and appended to the trace when the thread is done: where the time inside a kernel goes
struct TraceMarks {
  unsigned long long t[64];
  int k[64];
  int n = 0;
  __device__ __forceinline__ void mark(int kind) {
    if (n < 64) {
      t[n] = trace_now();
      k[n] = kind;
      ++n;
    }
  }
  __device__ __forceinline__ void flush() {
    if (blockIdx.x != 0 || blockIdx.y != 0 || !g_trace.rec || n == 0) return;
    const unsigned int i0 = atomicAdd(g_trace.cursor, (unsigned int)n);
    for (int j = 0; j < n; ++j) {
      if (i0 + j >= g_trace.cap) break;
      g_trace.rec[4 * (i0 + j) + 0] = (unsigned long long)k[j];
      g_trace.rec[4 * (i0 + j) + 1] = g_trace.rec[4 * (i0 + j) + 2] = g_trace.rec[4 * (i0 + j) + 3] = t[j];
    }
  }
};
#define VFM_TRACE_ATTACH(fn)                                                                       \
  extern "C" __attribute__((visibility("default"))) int fn(void* rec, void* cursor, unsigned cap) { \
    vfm::TraceBuf t{(unsigned long long*)rec, (unsigned int*)cursor, cap};                          \
    return (int)cudaMemcpyToSymbol(vfm::g_trace, &t, sizeof(t));                                    \
  }
#else
struct TraceScope {
  __device__ __forceinline__ explicit TraceScope(int) {}
  __device__ __forceinline__ void waited() {}
  __device__ __forceinline__ void end() {}
};
struct TraceMarks {
  __device__ __forceinline__ void mark(int) {}
  __device__ __forceinline__ void flush() {}
};
#define VFM_TRACE_ATTACH(fn)
#endif

// --- thread-block-cluster variants (CTA pair sharing the streamed operand by TMA multicast) ---
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
      ::"r"(dst), "l"(map), "r"(bar), "h"(cta_mask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// --- CTA-pair MMA (cta_group::2): one instruction issued by the leader CTA (cluster rank 0) drives the tensor cores of
// both SMs: M = 256 (128 accumulator rows in each CTA's TMEM), each CTA supplies its own 128 rows of A and HALF of the
// N-wide B tile from its own shared memory (same offsets in both CTAs) ---
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
  return r;
}
// arrive on a barrier that may live in the peer CTA (address from mapa_cluster).  Default semantics (release at CTA scope),
// as CUTLASS' ClusterBarrier::arrive: what the barrier orders here is TMEM reads, which tcgen05.wait::ld +
// tcgen05.fence::before_thread_sync already cover.  (`.release.cluster` compiled to MEMBAR.ALL.GPU + ERRBAR per arrive and
// cost the epilogue warps a fifth of their time -- ncu source page, profiles/r1_kernel_bench.txt.)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// TMA load into this CTA's shared memory whose completion bytes are credited to a barrier in the leader CTA
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}
// the same load multicast to the CTAs of `cta_mask` (same CTA-relative destination offset in each; every destination CTA's
// bytes are credited to the barrier at this offset in ITS pair's leader when `cluster_bar` points at the issuer's pair leader)
__device__ __forceinline__ void tma_load_2d_2sm_mc(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(dst),
      "l"(map), "r"(cluster_bar), "h"(cta_mask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
// completion of the leader's earlier cta_group::2 MMAs -> arrive on the barrier at this offset in the CTAs of `cta_mask`
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}
// pair-wide TMEM allocation: the same warp of BOTH CTAs issues it with the same destination offset
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}

// TMEM allocation (one warp, power-of-two columns >= 32) -- address is written to shared memory
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}

// instruction descriptor for kind::f16: fp32 accumulate, K-major A and B; ab_fmt 0 = fp16, 1 = bf16
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n, int ab_fmt) {
  return (1u << 4) | ((uint32_t)ab_fmt << 7) | ((uint32_t)ab_fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// 2-D row-major 16-bit tensor (rows x cols, row pitch in elements), box = 64 columns (128 B) x box_rows, 128B swizzle
int make_tmap_16bit(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t pitch_elems, int box_rows, bool bf16);

}  // namespace vfm
