// Micro-benchmark: tcgen05.ld shapes (32x32b .x16/.x32/.x64) on an otherwise idle SM -- cycles per instruction and bytes/cycle
// per warp.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/_bin/tmem_ld_bench2 tools/micro/tmem_ld_bench2.cu
// Only two of the loaded registers are consumed per load, so the loop is paced by the load itself.  (A first version
// xor-ed all N registers into one accumulator and measured that dependent chain instead: 2.7 cycles per register, i.e. the
// "85 cycles per .x32 load" once quoted.  The number that matters for the match epilogue -- ~245 cycles per .x32 load and SM
// sub-partition WHILE tcgen05.mma runs -- comes from the kernel's own counters, profiles/r1_kernel_bench.txt.)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int N> struct Ld;
#define REGS16(o) "=r"(r[o+0]), "=r"(r[o+1]), "=r"(r[o+2]), "=r"(r[o+3]), "=r"(r[o+4]), "=r"(r[o+5]), "=r"(r[o+6]), "=r"(r[o+7]), \
                  "=r"(r[o+8]), "=r"(r[o+9]), "=r"(r[o+10]), "=r"(r[o+11]), "=r"(r[o+12]), "=r"(r[o+13]), "=r"(r[o+14]), "=r"(r[o+15])
template <> struct Ld<16> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : REGS16(0) : "r"(a));
  }
};
template <> struct Ld<32> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : REGS16(0), REGS16(16) : "r"(a));
  }
};
template <> struct Ld<64> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                 "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                 : REGS16(0), REGS16(16), REGS16(32), REGS16(48) : "r"(a));
  }
};

template <int N>
__global__ void __launch_bounds__(256, 1) bench(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  uint32_t r[N];
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    Ld<N>::ld(base + (uint32_t)((i * N) & 511 & ~(N - 1)), r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc ^= r[0] ^ r[N - 1];
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
  sink[threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u));
}

template <int N>
void run(int warps, long long* d_out, uint32_t* d_sink) {
  const int iters = 4096;
  bench<N><<<1, warps * 32>>>(iters, d_out, d_sink);
  bench<N><<<1, warps * 32>>>(iters, d_out, d_sink);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return; }
  long long h[8];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
  const double c = (double)mx / iters;
  printf("32x32b.x%-3d warps=%d: %.1f cycles per load, %.1f B/cycle per warp, %.1f B/cycle/SM\n", N, warps, c, 128.0 * N / c, 128.0 * N * warps / c);
}

int main() {
  long long* d_out;
  uint32_t* d_sink;
  cudaMalloc(&d_out, 64 * sizeof(long long));
  cudaMalloc(&d_sink, 4096 * sizeof(uint32_t));
  for (int warps : {1, 4}) {
    run<16>(warps, d_out, d_sink);
    run<32>(warps, d_out, d_sink);
    run<64>(warps, d_out, d_sink);
  }
  return 0;
}
