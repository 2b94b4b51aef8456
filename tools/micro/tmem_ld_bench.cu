// Micro-benchmark: TMEM read throughput of tcgen05.ld on sm_100a (how fast can an epilogue drain an accumulator tile?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/_bin/tmem_ld_bench tools/micro/tmem_ld_bench.cu
// Each of W warps (W = 1, 2, 4, 8; warp w reads TMEM lane quarter w % 4) issues ITERS x tcgen05.ld.32x32b.x32 (4 KB per
// warp-instruction) and waits; reports cycles per instruction and bytes/cycle/SM.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

template <int INFLIGHT>
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
  uint32_t r[INFLIGHT][32];
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i += INFLIGHT) {
#pragma unroll
    for (int u = 0; u < INFLIGHT; ++u) ld32(base + (uint32_t)(((i + u) * 32) & 511), r[u]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int u = 0; u < INFLIGHT; ++u)
#pragma unroll
      for (int k = 0; k < 32; ++k) acc ^= r[u][k];
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[blockIdx.x * 8 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u));
}

template <int INFLIGHT>
void run(int warps, int iters, long long* d_out, uint32_t* d_sink) {
  bench<INFLIGHT><<<1, warps * 32>>>(iters, d_out, d_sink);
  bench<INFLIGHT><<<1, warps * 32>>>(iters, d_out, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
  long long h[8];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
  const double cyc_per = (double)mx / iters;
  printf("warps=%d inflight=%d: %.1f cycles per x32 load per warp, %.1f B/cycle/SM\n", warps, INFLIGHT, cyc_per,
         4096.0 * warps / cyc_per);
}

int main() {
  long long* d_out;
  uint32_t* d_sink;
  cudaMalloc(&d_out, 64 * sizeof(long long));
  cudaMalloc(&d_sink, 4096 * sizeof(uint32_t));
  for (int warps : {1, 2, 4, 8}) {
    run<1>(warps, 4096, d_out, d_sink);
    run<2>(warps, 4096, d_out, d_sink);
  }
  return 0;
}
