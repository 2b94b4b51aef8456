// Host -> device copies of caller-owned buffers.
//
// The reference's call surface hands over NumPy arrays, i.e. PAGEABLE host memory.  cudaMemcpyAsync from pageable memory is
// staged by the driver through its own pinned buffer by one thread: ~10 GB/s on the B200 boxes, against 55 GB/s from pinned
// memory (tools/h2d_bw.py) -- the whole end-to-end rate of register_batch() on plain NumPy inputs (31 MB per pair).
// h2d_copy() keeps the 55 GB/s for pinned / registered sources (one cudaMemcpyAsync) and does the staging itself otherwise: a
// small pool of host threads copies the source, chunk by chunk, into a ring of pinned staging buffers, each chunk going out by
// DMA while the next ones are being filled.  The call returns when the last chunk is ENQUEUED; the source may be reused as
// soon as it returns (its bytes have all been read), the destination is ready in stream order.
#include <sched.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace vfm {

constexpr size_t HC_CHUNK = 8u << 20;   // 8 MB per staging slot
constexpr int HC_SLOTS = 4;

struct HostCopy {
  char* ring = nullptr;                 // HC_SLOTS x HC_CHUNK pinned bytes
  cudaEvent_t done[HC_SLOTS] = {};      // the slot's DMA has completed
  bool used[HC_SLOTS] = {};
  int next = 0;
  // fork-join pool: the caller fills part 0, worker w fills part w + 1
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  uint64_t epoch = 0;
  int pending = 0;
  bool quit = false;
  char* job_dst = nullptr;
  const char* job_src = nullptr;
  size_t job_bytes = 0;
  int parts = 1;

  static void part_range(size_t bytes, int parts, int p, size_t* lo, size_t* hi) {
    const size_t per = ((bytes + parts - 1) / parts + 4095) & ~size_t(4095);
    *lo = per * p < bytes ? per * p : bytes;
    *hi = per * (p + 1) < bytes ? per * (p + 1) : bytes;
  }
  void worker_main(int w) {
    uint64_t seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> lk(mu);
      cv_go.wait(lk, [&] { return quit || epoch != seen; });
      if (quit) return;
      seen = epoch;
      char* d = job_dst;
      const char* s = job_src;
      const size_t n = job_bytes;
      const int np = parts;
      lk.unlock();
      size_t lo, hi;
      part_range(n, np, w + 1, &lo, &hi);
      if (hi > lo) memcpy(d + lo, s + lo, hi - lo);
      lk.lock();
      if (--pending == 0) cv_done.notify_one();
    }
  }
  void fill(char* dst, const char* src, size_t bytes) {   // parallel memcpy, returns when every part is done
    if (workers.empty() || bytes < (1u << 20)) {
      memcpy(dst, src, bytes);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(mu);
      job_dst = dst;
      job_src = src;
      job_bytes = bytes;
      pending = (int)workers.size();
      ++epoch;
    }
    cv_go.notify_all();
    size_t lo, hi;
    part_range(bytes, parts, 0, &lo, &hi);
    if (hi > lo) memcpy(dst + lo, src + lo, hi - lo);
    std::unique_lock<std::mutex> lk(mu);
    cv_done.wait(lk, [&] { return pending == 0; });
  }
  ~HostCopy() {
    {
      std::lock_guard<std::mutex> lk(mu);
      quit = true;
    }
    cv_go.notify_all();
    for (auto& t : workers) t.join();
    for (int s = 0; s < HC_SLOTS; ++s)
      if (done[s]) cudaEventDestroy(done[s]);
    if (ring) cudaFreeHost(ring);
  }
};

void hostcopy_destroy(vfmreg_ctx* ctx) {
  delete ctx->hostcopy;
  ctx->hostcopy = nullptr;
}

static int hostcopy_get(vfmreg_ctx* ctx, HostCopy** out) {
  if (!ctx->hostcopy) {
    HostCopy* h = new HostCopy();
    cudaError_t e = cudaMallocHost(&h->ring, HC_SLOTS * HC_CHUNK);
    if (e != cudaSuccess) {
      delete h;
      set_error("h2d_copy: cudaMallocHost(%zu) failed: %s", HC_SLOTS * HC_CHUNK, cudaGetErrorString(e));
      return VFMREG_ERR_ALLOC;
    }
    for (int s = 0; s < HC_SLOTS; ++s) cudaEventCreateWithFlags(&h->done[s], cudaEventDisableTiming);
    // copy threads: VFMREG_COPY_THREADS, default = half of the cores this process may run on (its affinity mask: one rank per
    // GPU under torchrun shares the box's cores), at most 8
    int threads = 0;
    if (const char* env = getenv("VFMREG_COPY_THREADS")) threads = atoi(env);
    if (threads <= 0) {
      cpu_set_t set;
      CPU_ZERO(&set);
      const int allowed = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
      threads = allowed / 2;
      threads = threads < 1 ? 1 : (threads > 8 ? 8 : threads);
    }
    h->parts = threads;
    for (int w = 0; w + 1 < threads; ++w) h->workers.emplace_back([h, w] { h->worker_main(w); });
    ctx->hostcopy = h;
  }
  *out = ctx->hostcopy;
  return VFMREG_OK;
}

int h2d_copy(vfmreg_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return VFMREG_OK;
  cudaPointerAttributes attr;
  const cudaError_t q = cudaPointerGetAttributes(&attr, src);
  if (q != cudaSuccess) cudaGetLastError();
  const bool pageable = (q != cudaSuccess) || attr.type == cudaMemoryTypeUnregistered;
  if (!pageable || bytes < (256u << 10)) {   // pinned / registered / managed, or too small to be worth a staging round
    VFM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    return VFMREG_OK;
  }
  HostCopy* h = nullptr;
  VFM_TRY(hostcopy_get(ctx, &h));
  for (size_t off = 0; off < bytes; off += HC_CHUNK) {
    const size_t len = bytes - off < HC_CHUNK ? bytes - off : HC_CHUNK;
    const int s = h->next;
    h->next = (h->next + 1) % HC_SLOTS;
    if (h->used[s]) VFM_CUDA(cudaEventSynchronize(h->done[s]));   // the slot's previous chunk has left the host
    h->fill(h->ring + (size_t)s * HC_CHUNK, static_cast<const char*>(src) + off, len);
    VFM_CUDA(cudaMemcpyAsync(static_cast<char*>(dst) + off, h->ring + (size_t)s * HC_CHUNK, len, cudaMemcpyHostToDevice, stream));
    VFM_CUDA(cudaEventRecord(h->done[s], stream));
    h->used[s] = true;
  }
  return VFMREG_OK;
}

}  // namespace vfm
