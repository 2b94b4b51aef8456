// Voxel hashing on the GPU: the steps either side of the match-and-solve path.
//
//   voxel_downsample   kiss_icp::VoxelDownsample        (Preprocessing.cpp:50-137): first point of every voxel, key =
//                      (p / voxel_size).cast<int>() -- truncation toward zero, computed in double as the binding does.
//   voxel map build    VoxelHashMap::AddPoints          (VoxelHashMap.cpp:735-771) + VoxelBlock::AddPoint: the first
//                      `max_points_per_voxel` points of every voxel, in insertion order.
//   nearest / ICP      VoxelHashMap::GetCorrespondences (VoxelHashMap.cpp:76-168: closest point among the 27 neighbouring
//                      voxels, voxels enumerated i, j, k ascending, strict '<' keeps the first minimum) and
//                      RegisterFrame / BuildLinearSystem (Registration.cpp:96-195): Gauss-Newton on point-to-point
//                      residuals with the Geman-McClure weight, dx = -(J^T W J)^-1 J^T W r, SE(3) exponential update,
//                      until |dx| < 1e-4.
//
// The reference's containers are tsl::robin_map; their ITERATION order (hence the order of VoxelDownsample's output and
// of VoxelHashMap::Pointcloud()) is an implementation detail of that hash table.  Here every output is ordered by the
// original point index, which makes the result a deterministic function of the input; the SET of points is the
// reference's.  All arithmetic is float64 with explicit operation order (file compiled with -fmad=false).
//
// Layout: open-addressing table of 64-bit keys (3 x 21-bit biased voxel coordinates), capacity = power of two >= 2n,
// linear probing; per slot the smallest point index (down-sampling) or a CSR bucket of kept point indices (map).
// Bound: HBM / L2 latency (12-24 B read per point, a handful of atomics); the ICP iteration is latency-bound
// (two launches per iteration, ~10^4 points).
#include <math.h>

#include "common.cuh"

struct vfmreg_voxel_map {
  double voxel_size = 1.0;
  int max_points = 20;
  int64_t n_points = 0;        // kept points
  uint32_t cap = 0;            // table capacity (power of two)
  unsigned long long* keys = nullptr;   // [cap]
  int32_t* start = nullptr;    // [cap] first kept point of the slot
  int32_t* num = nullptr;      // [cap] kept points of the slot (<= max_points)
  double* pts = nullptr;       // [n_points][3], grouped by slot, ascending source index inside a slot
  int32_t* src_idx = nullptr;  // [n_points] index into the array given to build()
  int device = 0;
};

namespace vfm {

constexpr unsigned long long EMPTY_KEY = ~0ULL;
constexpr int COORD_BIAS = 1 << 20;

__host__ __device__ __forceinline__ uint32_t hash_key(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return (uint32_t)k;
}

// (p / voxel_size).cast<int>() per coordinate; false when a coordinate leaves the 21-bit range (or is not finite)
__device__ __forceinline__ bool voxel_of(double x, double y, double z, double vs, int& ix, int& iy, int& iz) {
  const double fx = x / vs, fy = y / vs, fz = z / vs;
  const double lim = (double)(COORD_BIAS - 2);
  if (!(fabs(fx) < lim) || !(fabs(fy) < lim) || !(fabs(fz) < lim)) return false;
  ix = __double2int_rz(fx);
  iy = __double2int_rz(fy);
  iz = __double2int_rz(fz);
  return true;
}

__device__ __forceinline__ unsigned long long pack_key(int ix, int iy, int iz) {
  return ((unsigned long long)(uint32_t)(ix + COORD_BIAS) << 42) | ((unsigned long long)(uint32_t)(iy + COORD_BIAS) << 21) |
         (unsigned long long)(uint32_t)(iz + COORD_BIAS);
}

// claim-or-find the slot of `key`
__device__ __forceinline__ uint32_t table_insert(unsigned long long* keys, uint32_t mask, unsigned long long key) {
  uint32_t s = hash_key(key) & mask;
  while (true) {
    const unsigned long long prev = atomicCAS(&keys[s], EMPTY_KEY, key);
    if (prev == EMPTY_KEY || prev == key) return s;
    s = (s + 1) & mask;
  }
}

// slot of `key`, or -1
__device__ __forceinline__ int table_find(const unsigned long long* __restrict__ keys, uint32_t mask, unsigned long long key) {
  uint32_t s = hash_key(key) & mask;
  while (true) {
    const unsigned long long k = __ldg(keys + s);
    if (k == key) return (int)s;
    if (k == EMPTY_KEY) return -1;
    s = (s + 1) & mask;
  }
}

template <typename T>
__device__ __forceinline__ void load_xyz(const void* pts, int64_t i, int cols, double& x, double& y, double& z) {
  const T* p = static_cast<const T*>(pts) + i * cols;
  x = (double)p[0];
  y = (double)p[1];
  z = (double)p[2];
}

// pass 1: every point finds / claims the slot of its voxel; first[slot] = min index, cnt[slot] += 1
template <typename T>
__global__ void __launch_bounds__(256) voxel_insert_kernel(const void* __restrict__ pts, int n, int cols, double vs,
                                                           unsigned long long* keys, uint32_t mask, int32_t* first,
                                                           int32_t* cnt, int32_t* slot_of, int32_t* bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x, y, z;
  load_xyz<T>(pts, i, cols, x, y, z);
  int ix, iy, iz;
  if (!voxel_of(x, y, z, vs, ix, iy, iz)) {
    atomicAdd(bad, 1);
    slot_of[i] = -1;
    return;
  }
  const uint32_t s = table_insert(keys, mask, pack_key(ix, iy, iz));
  slot_of[i] = (int32_t)s;
  if (first) atomicMin(&first[s], i);
  if (cnt) atomicAdd(&cnt[s], 1);
}

// ---- ordered stream compaction over n flags: block counts -> scan -> scatter -------------------------------------
constexpr int CB = 1024;

__global__ void __launch_bounds__(CB) flag_first_kernel(const int32_t* __restrict__ slot_of, const int32_t* __restrict__ first, int n,
                                                       uint8_t* __restrict__ flag, int32_t* __restrict__ block_count) {
  __shared__ int wsum[32];
  const int i = blockIdx.x * CB + threadIdx.x;
  bool keep = false;
  if (i < n) {
    const int s = slot_of[i];
    keep = s >= 0 && first[s] == i;
    flag[i] = keep ? 1 : 0;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = __popc(bal);
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = wsum[threadIdx.x];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (threadIdx.x == 0) block_count[blockIdx.x] = v;
  }
}

// exclusive scan of `m` ints by one CTA (m up to a few 10^5: the table or the block counts); total -> *total_out
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int m,
                                                              int32_t* __restrict__ total_out) {
  __shared__ int wsum[32];
  __shared__ int carry_s;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < m; base += 1024 * 4) {
    int v[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = base + t * 4 + k;
      v[k] = (i < m) ? in[i] : 0;
      sum += v[k];
    }
    int incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += u;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
      int x = wsum[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += u;
      }
      wsum[lane] = x;   // inclusive over warps
    }
    __syncthreads();
    const int carry = carry_s;
    int run = carry + (w > 0 ? wsum[w - 1] : 0) + incl - sum;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = base + t * 4 + k;
      if (i < m) out[i] = run;
      run += v[k];
    }
    __syncthreads();
    if (t == 1023) carry_s = carry + wsum[31];
    __syncthreads();
  }
  if (t == 0 && total_out) *total_out = carry_s;
}

__global__ void __launch_bounds__(CB) scatter_flagged_kernel(const uint8_t* __restrict__ flag, const int32_t* __restrict__ block_off, int n,
                                                            int32_t* __restrict__ out_idx) {
  __shared__ int wsum[32];
  const int i = blockIdx.x * CB + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool keep = (i < n) && flag[i];
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) wsum[w] = __popc(bal);
  __syncthreads();
  int off = block_off[blockIdx.x];
  for (int k = 0; k < w; ++k) off += wsum[k];
  if (keep) out_idx[off + __popc(bal & ((1u << lane) - 1u))] = i;
}

// out[k] = rows[idx[k]] for k < *count; rows of row_bytes (multiple of 4) bytes; one warp per row
__global__ void __launch_bounds__(256) gather_rows_generic_kernel(const uint32_t* __restrict__ rows, int row_words,
                                                                 const int32_t* __restrict__ idx, const int32_t* __restrict__ count,
                                                                 int max_rows, uint32_t* __restrict__ out) {
  const int rows_n = min(*count, max_rows);
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < rows_n; k += warps) {
    const uint32_t* s = rows + (size_t)idx[k] * row_words;
    uint32_t* d = out + (size_t)k * row_words;
    for (int q = lane; q < row_words; q += 32) d[q] = __ldg(s + q);
  }
}

// ---- voxel map build -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bucket_scatter_kernel(const int32_t* __restrict__ slot_of, int n, const int32_t* __restrict__ off,
                                                            int32_t* __restrict__ cursor, int32_t* __restrict__ bucket) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = slot_of[i];
  if (s < 0) return;
  bucket[off[s] + atomicAdd(&cursor[s], 1)] = i;
}

// one thread per slot: kept[slot] = min(cnt, max_points); the kept indices are the max_points smallest of the bucket,
// moved to its front in ascending order (selection by repeated minimum: buckets are small)
__global__ void __launch_bounds__(128) bucket_select_kernel(const int32_t* __restrict__ off, const int32_t* __restrict__ cnt, uint32_t cap,
                                                           int max_points, int32_t* __restrict__ bucket, int32_t* __restrict__ kept) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= cap) return;
  const int c = cnt[s];
  const int keep = min(c, max_points);
  kept[s] = keep;
  int32_t* b = bucket + off[s];
  for (int k = 0; k < keep; ++k) {
    int best = k;
    int bv = b[k];
    for (int j = k + 1; j < c; ++j) {
      const int v = b[j];
      if (v < bv) {
        bv = v;
        best = j;
      }
    }
    b[best] = b[k];
    b[k] = bv;
  }
}

__global__ void __launch_bounds__(128) map_fill_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ off,
                                                      const int32_t* __restrict__ bucket, const int32_t* __restrict__ kept,
                                                      const int32_t* __restrict__ start, uint32_t cap, double* __restrict__ pts,
                                                      int32_t* __restrict__ src_idx) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= cap) return;
  const int keep = kept[s];
  const int32_t* b = bucket + off[s];
  const int o = start[s];
  for (int k = 0; k < keep; ++k) {
    const int i = b[k];
    pts[(size_t)(o + k) * 3 + 0] = xyz[(size_t)i * 3 + 0];
    pts[(size_t)(o + k) * 3 + 1] = xyz[(size_t)i * 3 + 1];
    pts[(size_t)(o + k) * 3 + 2] = xyz[(size_t)i * 3 + 2];
    src_idx[o + k] = i;
  }
}

// ---- nearest neighbour among the 27 voxels around a point (VoxelHashMap.cpp:79-136) --------------------------------
struct MapView {
  const unsigned long long* keys;
  const int32_t* start;
  const int32_t* num;
  const double* pts;
  uint32_t mask;
  double vs;
};

// One warp per query: lane l < 27 searches voxel l of the (i, j, k)-ascending enumeration (l = 9 (i - kx + 1) + 3 (j - ky + 1)
// + (k - kz + 1)), strict '<' inside a voxel keeps the first minimum; the warp-wide arg-min prefers the lower lane on equal
// distances, i.e. the earlier voxel -- the same winner as the reference's sequential scan.  All lanes return the result.
__device__ __forceinline__ int nearest_in_map(const MapView& M, double x, double y, double z, double& best_d2) {
  const int lane = threadIdx.x & 31;
  // static_cast<int>(point[k] / voxel_size_): truncation toward zero
  const int kx = __double2int_rz(x / M.vs), ky = __double2int_rz(y / M.vs), kz = __double2int_rz(z / M.vs);
  int best = -1;
  double bd = 1.7976931348623157e308;   // std::numeric_limits<double>::max()
  if (lane < 27) {
    const int i = kx - 1 + lane / 9, j = ky - 1 + (lane / 3) % 3, k = kz - 1 + lane % 3;
    if (abs(i) < COORD_BIAS - 2 && abs(j) < COORD_BIAS - 2 && abs(k) < COORD_BIAS - 2) {
      const int s = table_find(M.keys, M.mask, pack_key(i, j, k));
      if (s >= 0) {
        const int o = M.start[s], c = M.num[s];
        for (int e = 0; e < c; ++e) {
          const double dx = M.pts[(size_t)(o + e) * 3 + 0] - x, dy = M.pts[(size_t)(o + e) * 3 + 1] - y,
                       dz = M.pts[(size_t)(o + e) * 3 + 2] - z;
          const double d2 = (dx * dx + dy * dy) + dz * dz;   // Eigen squaredNorm of a 3-vector, no contraction
          if (d2 < bd) {
            bd = d2;
            best = o + e;
          }
        }
      }
    }
  }
  int who = lane;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, bd, off);
    const int ob = __shfl_xor_sync(0xffffffffu, best, off);
    const int ow = __shfl_xor_sync(0xffffffffu, who, off);
    if (od < bd || (od == bd && ow < who)) {
      bd = od;
      best = ob;
      who = ow;
    }
  }
  best_d2 = bd;
  return best;
}

__global__ void __launch_bounds__(128) map_nearest_kernel(MapView M, const double* __restrict__ q, int n, double max_dist,
                                                         int32_t* __restrict__ nn, double* __restrict__ nn_d2) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // warp per query
  if (i >= n) return;
  double d2;
  const int b = nearest_in_map(M, q[(size_t)i * 3], q[(size_t)i * 3 + 1], q[(size_t)i * 3 + 2], d2);
  if ((threadIdx.x & 31) != 0) return;
  const bool ok = b >= 0 && sqrt(d2) < max_dist;   // (closest - point).norm() < max_correspondance_distance
  nn[i] = ok ? b : -1;
  if (nn_d2) nn_d2[i] = b >= 0 ? d2 : -1.0;
}

// ---- ICP (Registration.cpp:145-195) --------------------------------------------------------------------------------
constexpr int ICP_THREADS = 128;
constexpr int ICP_TERMS = 28;   // 21 upper-triangular J^T W J entries + 6 J^T W r entries + correspondence count

struct IcpState {          // device
  double est[12];          // pending update (R row-major, t): applied to the source points by the next correspondence pass
  double T[12];            // T_icp * initial_guess so far
  int32_t done;            // 1: converged / no correspondences
  int32_t iters;           // iterations executed (solves)
  int32_t last_corr;       // correspondences of the last iteration
  int32_t pad;
};

// one pass: source <- est * source (the previous iteration's update, Registration.cpp:177), nearest neighbours (one warp
// per point, points strided over a fixed grid), per-CTA partial sums of the normal equations in a fixed order
// (deterministic for a given grid size)
__global__ void __launch_bounds__(ICP_THREADS) icp_accumulate_kernel(MapView M, double* __restrict__ src, int n, double max_dist,
                                                                    double kernel, const IcpState* __restrict__ st,
                                                                    double* __restrict__ partial) {
  __shared__ double red[ICP_THREADS / 32][ICP_TERMS];
  if (st->done) return;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int warps = (gridDim.x * ICP_THREADS) >> 5;
  double acc[ICP_TERMS];
#pragma unroll
  for (int k = 0; k < ICP_TERMS; ++k) acc[k] = 0.0;
  double e[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) e[k] = st->est[k];
  for (int i = (blockIdx.x * ICP_THREADS + threadIdx.x) >> 5; i < n; i += warps) {
    const double px = src[(size_t)i * 3], py = src[(size_t)i * 3 + 1], pz = src[(size_t)i * 3 + 2];
    const double x = ((e[0] * px + e[1] * py) + e[2] * pz) + e[9];
    const double y = ((e[3] * px + e[4] * py) + e[5] * pz) + e[10];
    const double z = ((e[6] * px + e[7] * py) + e[8] * pz) + e[11];
    double d2;
    const int b = nearest_in_map(M, x, y, z, d2);
    __syncwarp();
    if (lane == 0) {
      src[(size_t)i * 3] = x;
      src[(size_t)i * 3 + 1] = y;
      src[(size_t)i * 3 + 2] = z;
      if (b >= 0 && sqrt(d2) < max_dist) {
        const double rx = x - M.pts[(size_t)b * 3], ry = y - M.pts[(size_t)b * 3 + 1], rz = z - M.pts[(size_t)b * 3 + 2];
        const double r2 = (rx * rx + ry * ry) + rz * rz;
        const double kk = kernel + r2;
        const double wgt = (kernel * kernel) / (kk * kk);   // square(kernel) / square(kernel + residual2)
        // J = [I | -hat(s)]: columns c3 = (0, -z, y), c4 = (z, 0, -x), c5 = (-y, x, 0)
        const double c[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, -z, y}, {z, 0, -x}, {-y, x, 0}};
        int t = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int bb = a; bb < 6; ++bb) acc[t++] += wgt * ((c[a][0] * c[bb][0] + c[a][1] * c[bb][1]) + c[a][2] * c[bb][2]);
#pragma unroll
        for (int a = 0; a < 6; ++a) acc[21 + a] += wgt * ((c[a][0] * rx + c[a][1] * ry) + c[a][2] * rz);
        acc[27] += 1.0;
      }
    }
  }
  if (lane == 0)
    for (int k = 0; k < ICP_TERMS; ++k) red[w][k] = acc[k];
  __syncthreads();
  if (threadIdx.x < ICP_TERMS) {
    double v = 0.0;
    for (int ww = 0; ww < ICP_THREADS / 32; ++ww) v += red[ww][threadIdx.x];
    partial[(size_t)blockIdx.x * ICP_TERMS + threadIdx.x] = v;
  }
}

// SE3::exp(dx), dx = (upsilon, omega): R = exp(hat(omega)) (Rodrigues), t = V upsilon
__device__ void se3_exp(const double* dx, double* rt) {
  const double ux = dx[0], uy = dx[1], uz = dx[2], wx = dx[3], wy = dx[4], wz = dx[5];
  const double th2 = (wx * wx + wy * wy) + wz * wz;
  const double th = sqrt(th2);
  double A, B, C;   // sin(th)/th, (1-cos(th))/th^2, (th - sin(th))/th^3
  if (th < 1e-5) {
    A = 1.0 - th2 / 6.0;
    B = 0.5 - th2 / 24.0;
    C = 1.0 / 6.0 - th2 / 120.0;
  } else {
    A = sin(th) / th;
    B = (1.0 - cos(th)) / th2;
    C = (th - sin(th)) / (th2 * th);
  }
  const double W[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
  double W2[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) W2[i * 3 + j] = (W[i * 3] * W[j] + W[i * 3 + 1] * W[3 + j]) + W[i * 3 + 2] * W[6 + j];
  double V[9];
  for (int i = 0; i < 9; ++i) {
    const double id = (i % 4 == 0) ? 1.0 : 0.0;
    rt[i] = (id + A * W[i]) + B * W2[i];
    V[i] = (id + B * W[i]) + C * W2[i];
  }
  rt[9] = (V[0] * ux + V[1] * uy) + V[2] * uz;
  rt[10] = (V[3] * ux + V[4] * uy) + V[5] * uz;
  rt[11] = (V[6] * ux + V[7] * uy) + V[8] * uz;
}

// c = a * b for (R, t) pairs
__device__ void rt_mul(const double* a, const double* b, double* c) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) c[i * 3 + j] = (a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j]) + a[i * 3 + 2] * b[6 + j];
    c[9 + i] = ((a[i * 3] * b[9] + a[i * 3 + 1] * b[10]) + a[i * 3 + 2] * b[11]) + a[9 + i];
  }
}

// dx = -(J^T W J)^-1 J^T W r from the 27 accumulated terms (LDL^T without pivoting: J^T W J is symmetric positive
// semi-definite), est = SE3::exp(dx); returns |dx|
__device__ double solve_update(const double* s, double* est) {
  double Amat[6][6], rhs[6];
  int t = 0;
  for (int a = 0; a < 6; ++a)
    for (int b = a; b < 6; ++b) {
      Amat[a][b] = s[t];
      Amat[b][a] = s[t];
      ++t;
    }
  for (int a = 0; a < 6; ++a) rhs[a] = -s[21 + a];
  double L[6][6], D[6];
  for (int j = 0; j < 6; ++j) {
    double d = Amat[j][j];
    for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k] * D[k];
    D[j] = d;
    for (int i = j + 1; i < 6; ++i) {
      double v = Amat[i][j];
      for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k] * D[k];
      L[i][j] = (d != 0.0) ? v / d : 0.0;
    }
  }
  double yv[6], dx[6];
  for (int i = 0; i < 6; ++i) {
    double v = rhs[i];
    for (int k = 0; k < i; ++k) v -= L[i][k] * yv[k];
    yv[i] = v;
  }
  for (int i = 5; i >= 0; --i) {
    double v = (D[i] != 0.0) ? yv[i] / D[i] : 0.0;
    for (int k = i + 1; k < 6; ++k) v -= L[k][i] * dx[k];
    dx[i] = v;
  }
  se3_exp(dx, est);
  double nrm = 0.0;
  for (int i = 0; i < 6; ++i) nrm += dx[i] * dx[i];
  return sqrt(nrm);
}

// one thread: fixed-order sum of the partials, solve, pose update, termination test (Registration.cpp:172-183)
__global__ void icp_solve_kernel(const double* __restrict__ partial, int n_blocks, IcpState* st) {
  __shared__ double s[32];
  if (st->done) return;
  if (threadIdx.x < ICP_TERMS) {   // lane k sums term k over the blocks, in block order
    double v = 0.0;
    for (int b = 0; b < n_blocks; ++b) v += partial[(size_t)b * ICP_TERMS + threadIdx.x];
    s[threadIdx.x] = v;
  }
  __syncwarp();
  if (threadIdx.x != 0) return;
  st->last_corr = (int32_t)s[27];
  if (s[27] == 0.0) {   // "No correspondences found": leave the loop, pose unchanged
    st->done = 1;
    for (int i = 0; i < 12; ++i) st->est[i] = (i < 9 && i % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  double est[12], Tn[12];
  const double nrm = solve_update(s, est);
  rt_mul(est, st->T, Tn);
  for (int i = 0; i < 12; ++i) {
    st->est[i] = est[i];
    st->T[i] = Tn[i];
  }
  st->iters += 1;
  if (nrm < 1e-4) st->done = 1;   // ESTIMATION_THRESHOLD_ (the update has been applied to the pose)
}

// ---- VFM-ICP, first loop (Registration.cpp:197-329): Gauss-Newton on a FIXED list of descriptor correspondences, pruned
// after every update to |d - median| < 1.5 * 1.4826 * MAD, until the mean distance changes by < 0.01 m.  K is a few
// hundred (the source is voxelised at 5 m first), so the whole loop is one CTA; medians come from a bitonic sort.
constexpr int P1_THREADS = 1024;

struct Phase1Out {
  double T[12];     // T_icp of this loop * initial guess
  int32_t iters;    // value of j when the loop was left
  int32_t kept;     // correspondences left
};

__device__ void block_sum(double* vals, int n_terms, double (*red)[ICP_TERMS], double* out) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int k = 0; k < n_terms; ++k) {
    double v = vals[k];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) red[w][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < n_terms) {
    double v = 0.0;
    for (int ww = 0; ww < P1_THREADS / 32; ++ww) v += red[ww][threadIdx.x];
    out[threadIdx.x] = v;
  }
  __syncthreads();
}

// ascending bitonic sort of buf[0 .. n2) (n2 a power of two) by the whole CTA
__device__ void block_bitonic_sort(double* buf, int n2) {
  for (int size = 2; size <= n2; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < n2 / 2; i += P1_THREADS) {
        const int lo = (i / stride) * 2 * stride + (i % stride), hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const double a = buf[lo], b = buf[hi];
        if ((a > b) == up) {
          buf[lo] = b;
          buf[hi] = a;
        }
      }
      __syncthreads();
    }
}

// std::nth_element median with the even-size average (Registration.cpp:283-293)
__device__ __forceinline__ double median_sorted(const double* sorted, int k) {
  const int n = k / 2;
  return (k & 1) ? sorted[n] : (sorted[n] + sorted[n - 1]) / 2;
}

__global__ void __launch_bounds__(P1_THREADS) vfm_icp_phase1_kernel(const double* __restrict__ src_in, const double* __restrict__ tgt_in, int K,
                                                                   const double* __restrict__ T0, double kernel, int max_iters,
                                                                   double* src, double* tgt, double* src2, double* tgt2, double* dist,
                                                                   double* sortbuf, Phase1Out* out) {
  __shared__ double red[P1_THREADS / 32][ICP_TERMS];
  __shared__ double sums[ICP_TERMS];
  __shared__ double est_s[12], T_s[12];
  __shared__ double stat_s[4];   // mean, median, mad
  __shared__ int wsum[32];
  __shared__ int base_s;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  // Equation (9): source <- initial_guess * source
  if (t < 12) T_s[t] = (t < 9) ? T0[(t / 3) * 4 + (t % 3)] : T0[(t - 9) * 4 + 3];
  __syncthreads();
  double acc1[1] = {0.0};
  for (int k = t; k < K; k += P1_THREADS) {
    const double px = src_in[k * 3], py = src_in[k * 3 + 1], pz = src_in[k * 3 + 2];
    const double x = ((T_s[0] * px + T_s[1] * py) + T_s[2] * pz) + T_s[9];
    const double y = ((T_s[3] * px + T_s[4] * py) + T_s[5] * pz) + T_s[10];
    const double z = ((T_s[6] * px + T_s[7] * py) + T_s[8] * pz) + T_s[11];
    src[k * 3] = x; src[k * 3 + 1] = y; src[k * 3 + 2] = z;
    const double qx = tgt_in[k * 3], qy = tgt_in[k * 3 + 1], qz = tgt_in[k * 3 + 2];
    tgt[k * 3] = qx; tgt[k * 3 + 1] = qy; tgt[k * 3 + 2] = qz;
    const double dx = x - qx, dy = y - qy, dz = z - qz;
    acc1[0] += sqrt((dx * dx + dy * dy) + dz * dz);
  }
  block_sum(acc1, 1, red, sums);
  double prev = sums[0] / (double)K;   // NaN for K == 0, never used then
  int j = 0;
  for (; j < max_iters; ++j) {
    if (K == 0) break;   // "No correspondences found"
    double acc[ICP_TERMS];
#pragma unroll
    for (int i = 0; i < ICP_TERMS; ++i) acc[i] = 0.0;
    for (int k = t; k < K; k += P1_THREADS) {
      const double x = src[k * 3], y = src[k * 3 + 1], z = src[k * 3 + 2];
      const double rx = x - tgt[k * 3], ry = y - tgt[k * 3 + 1], rz = z - tgt[k * 3 + 2];
      const double r2 = (rx * rx + ry * ry) + rz * rz;
      const double kk = kernel + r2;
      const double wgt = (kernel * kernel) / (kk * kk);
      const double c[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, -z, y}, {z, 0, -x}, {-y, x, 0}};
      int q = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = a; b < 6; ++b) acc[q++] += wgt * ((c[a][0] * c[b][0] + c[a][1] * c[b][1]) + c[a][2] * c[b][2]);
#pragma unroll
      for (int a = 0; a < 6; ++a) acc[21 + a] += wgt * ((c[a][0] * rx + c[a][1] * ry) + c[a][2] * rz);
    }
    block_sum(acc, 27, red, sums);
    if (t == 0) {
      double est[12], Tn[12];
      solve_update(sums, est);
      rt_mul(est, T_s, Tn);
      for (int i = 0; i < 12; ++i) {
        est_s[i] = est[i];
        T_s[i] = Tn[i];
      }
    }
    __syncthreads();
    // TransformPoints(estimation, src_3d); distances; their mean
    int n2 = 1;
    while (n2 < K) n2 <<= 1;
    acc1[0] = 0.0;
    for (int k = t; k < n2; k += P1_THREADS) {
      double d = INFINITY;
      if (k < K) {
        const double px = src[k * 3], py = src[k * 3 + 1], pz = src[k * 3 + 2];
        const double x = ((est_s[0] * px + est_s[1] * py) + est_s[2] * pz) + est_s[9];
        const double y = ((est_s[3] * px + est_s[4] * py) + est_s[5] * pz) + est_s[10];
        const double z = ((est_s[6] * px + est_s[7] * py) + est_s[8] * pz) + est_s[11];
        src[k * 3] = x; src[k * 3 + 1] = y; src[k * 3 + 2] = z;
        const double dx = x - tgt[k * 3], dy = y - tgt[k * 3 + 1], dz = z - tgt[k * 3 + 2];
        d = sqrt((dx * dx + dy * dy) + dz * dz);
        dist[k] = d;
        acc1[0] += d;
      }
      sortbuf[k] = d;
    }
    block_sum(acc1, 1, red, sums);
    block_bitonic_sort(sortbuf, n2);
    if (t == 0) {
      stat_s[0] = sums[0] / (double)K;
      stat_s[1] = median_sorted(sortbuf, K);
    }
    __syncthreads();
    const double mean = stat_s[0], median = stat_s[1];
    for (int k = t; k < n2; k += P1_THREADS) sortbuf[k] = (k < K) ? fabs(dist[k] - median) : INFINITY;
    __syncthreads();
    block_bitonic_sort(sortbuf, n2);
    if (t == 0) stat_s[2] = median_sorted(sortbuf, K) * 1.4826;
    if (t == 0) base_s = 0;
    __syncthreads();
    const double mad = stat_s[2];
    // keep |d - median| < 1.5 * mad, order preserved
    for (int start = 0; start < K; start += P1_THREADS) {
      const int k = start + t;
      const bool keep = (k < K) && (fabs(dist[k] - median) < 1.5 * mad);
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) wsum[w] = __popc(bal);
      __syncthreads();
      int off = 0, tot = 0;
      for (int q = 0; q < 32; ++q) {
        const int vv = wsum[q];
        if (q < w) off += vv;
        tot += vv;
      }
      const int base = base_s;
      if (keep) {
        const int pos = base + off + __popc(bal & ((1u << lane) - 1u));
        for (int c = 0; c < 3; ++c) {
          src2[pos * 3 + c] = src[k * 3 + c];
          tgt2[pos * 3 + c] = tgt[k * 3 + c];
        }
      }
      __syncthreads();
      if (t == 0) base_s = base + tot;
      __syncthreads();
    }
    K = base_s;
    double* sw = src; src = src2; src2 = sw;
    sw = tgt; tgt = tgt2; tgt2 = sw;
    __syncthreads();
    if (fabs(prev - mean) < 0.01) break;   // EUCL_DIST_THRESHOLD_ (j is not advanced by the break)
    prev = mean;
  }
  if (t < 12) out->T[t] = T_s[t];
  if (t == 0) {
    out->iters = j;
    out->kept = K;
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------
static uint32_t table_capacity(int64_t n) {
  uint32_t cap = 1024;
  while ((int64_t)cap < 2 * n) cap <<= 1;
  return cap;
}

size_t voxel_scratch(int64_t n) {
  const uint32_t cap = table_capacity(n);
  return arena_bytes(cap, 8) + arena_bytes(cap, 4) * 6 + arena_bytes(n, 4) * 3 + arena_bytes(n, 1) + arena_bytes((size_t)ceil_div(n, CB) + 1, 4) * 2 +
         arena_bytes(4, 4) + 8192;
}

template <typename T>
static void launch_insert(vfmreg_ctx* ctx, const void* pts, int64_t n, int cols, double vs, unsigned long long* keys, uint32_t cap,
                          int32_t* first, int32_t* cnt, int32_t* slot_of, int32_t* bad) {
  voxel_insert_kernel<T><<<ceil_div(n, 256), 256, 0, ctx->stream>>>(pts, (int)n, cols, vs, keys, cap - 1, first, cnt, slot_of, bad);
}

int voxel_downsample(vfmreg_ctx* ctx, const void* pts, int64_t n, int cols, int elem_size, double voxel_size, int32_t* keep_idx,
                     int32_t* count) {
  VFM_CHECK_ARG(n > 0 && n < (1LL << 28), "voxel_downsample: point count %lld outside (0, 2^28)", (long long)n);
  VFM_CHECK_ARG(cols >= 3 && (elem_size == 4 || elem_size == 8), "voxel_downsample: need >= 3 columns of float32 / float64");
  VFM_CHECK_ARG(voxel_size > 0.0, "voxel_downsample: voxel_size must be > 0");
  const uint32_t cap = table_capacity(n);
  const int blocks = ceil_div(n, CB);
  unsigned long long* keys = arena_take<unsigned long long>(ctx, cap);
  int32_t* first = arena_take<int32_t>(ctx, cap);
  int32_t* slot_of = arena_take<int32_t>(ctx, n);
  uint8_t* flag = arena_take<uint8_t>(ctx, n);
  int32_t* bcount = arena_take<int32_t>(ctx, blocks + 1);
  int32_t* boff = arena_take<int32_t>(ctx, blocks + 1);
  int32_t* bad = arena_take<int32_t>(ctx, 4);
  if (!keys || !first || !slot_of || !flag || !bcount || !boff || !bad) {
    set_error("voxel_downsample: scratch arena too small");
    return VFMREG_ERR_ALLOC;
  }
  VFM_CUDA(cudaMemsetAsync(keys, 0xFF, sizeof(unsigned long long) * cap, ctx->stream));
  VFM_CUDA(cudaMemsetAsync(first, 0x7F, sizeof(int32_t) * cap, ctx->stream));
  VFM_CUDA(cudaMemsetAsync(bad, 0, sizeof(int32_t) * 4, ctx->stream));
  if (elem_size == 4)
    launch_insert<float>(ctx, pts, n, cols, voxel_size, keys, cap, first, nullptr, slot_of, bad);
  else
    launch_insert<double>(ctx, pts, n, cols, voxel_size, keys, cap, first, nullptr, slot_of, bad);
  VFM_TRY(launch_check(ctx, "voxel_insert_kernel"));
  flag_first_kernel<<<blocks, CB, 0, ctx->stream>>>(slot_of, first, (int)n, flag, bcount);
  VFM_TRY(launch_check(ctx, "flag_first_kernel"));
  exclusive_scan_kernel<<<1, 1024, 0, ctx->stream>>>(bcount, boff, blocks, count);
  VFM_TRY(launch_check(ctx, "exclusive_scan_kernel"));
  scatter_flagged_kernel<<<blocks, CB, 0, ctx->stream>>>(flag, boff, (int)n, keep_idx);
  VFM_TRY(launch_check(ctx, "scatter_flagged_kernel"));
  int32_t bad_h = 0;
  VFM_CUDA(cudaMemcpyAsync(&bad_h, bad, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  VFM_CUDA(cudaStreamSynchronize(ctx->stream));
  VFM_CHECK_ARG(bad_h == 0, "voxel_downsample: %d point(s) outside the +-2^20 voxel range or not finite", bad_h);
  return VFMREG_OK;
}

int gather_rows_generic(vfmreg_ctx* ctx, const void* rows, int row_bytes, const int32_t* idx, const int32_t* count, int64_t max_rows,
                        void* out) {
  VFM_CHECK_ARG(row_bytes > 0 && row_bytes % 4 == 0, "gather_rows: row size %d not a multiple of 4 bytes", row_bytes);
  if (max_rows <= 0) return VFMREG_OK;
  const int64_t want = (max_rows + 7) / 8;
  const int blocks = (int)(want < ctx->sm_count * 8 ? want : ctx->sm_count * 8);
  gather_rows_generic_kernel<<<blocks, 256, 0, ctx->stream>>>(static_cast<const uint32_t*>(rows), row_bytes / 4, idx, count,
                                                             (int)max_rows, static_cast<uint32_t*>(out));
  return launch_check(ctx, "gather_rows_generic_kernel");
}

static void map_free(vfmreg_voxel_map* m) {
  cudaFree(m->keys);
  cudaFree(m->start);
  cudaFree(m->num);
  cudaFree(m->pts);
  cudaFree(m->src_idx);
  m->keys = nullptr;
  m->start = m->num = m->src_idx = nullptr;
  m->pts = nullptr;
  m->n_points = 0;
  m->cap = 0;
}

int voxel_map_build(vfmreg_ctx* ctx, vfmreg_voxel_map* m, const double* xyz, int64_t n) {
  VFM_CHECK_ARG(m, "voxel_map_build: null map");
  VFM_CHECK_ARG(n >= 0 && n < (1LL << 28), "voxel_map_build: point count outside [0, 2^28)");
  VFM_CUDA(cudaStreamSynchronize(ctx->stream));
  map_free(m);
  if (n == 0) return VFMREG_OK;
  const uint32_t cap = table_capacity(n);
  int32_t *cnt = arena_take<int32_t>(ctx, cap), *off = arena_take<int32_t>(ctx, cap), *cursor = arena_take<int32_t>(ctx, cap);
  int32_t *kept = arena_take<int32_t>(ctx, cap), *slot_of = arena_take<int32_t>(ctx, n), *bucket = arena_take<int32_t>(ctx, n);
  int32_t* scal = arena_take<int32_t>(ctx, 4);   // [0] bad, [1] total bucketed, [2] total kept
  if (!cnt || !off || !cursor || !kept || !slot_of || !bucket || !scal) {
    set_error("voxel_map_build: scratch arena too small");
    return VFMREG_ERR_ALLOC;
  }
  VFM_CUDA(cudaMalloc(&m->keys, sizeof(unsigned long long) * cap));
  VFM_CUDA(cudaMalloc(&m->start, sizeof(int32_t) * cap));
  VFM_CUDA(cudaMalloc(&m->num, sizeof(int32_t) * cap));
  m->cap = cap;
  VFM_CUDA(cudaMemsetAsync(m->keys, 0xFF, sizeof(unsigned long long) * cap, ctx->stream));
  VFM_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * cap, ctx->stream));
  VFM_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * cap, ctx->stream));
  VFM_CUDA(cudaMemsetAsync(scal, 0, sizeof(int32_t) * 4, ctx->stream));
  launch_insert<double>(ctx, xyz, n, 3, m->voxel_size, m->keys, cap, nullptr, cnt, slot_of, scal);
  VFM_TRY(launch_check(ctx, "voxel_insert_kernel"));
  exclusive_scan_kernel<<<1, 1024, 0, ctx->stream>>>(cnt, off, (int)cap, scal + 1);
  VFM_TRY(launch_check(ctx, "exclusive_scan_kernel"));
  bucket_scatter_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(slot_of, (int)n, off, cursor, bucket);
  VFM_TRY(launch_check(ctx, "bucket_scatter_kernel"));
  bucket_select_kernel<<<ceil_div(cap, 128), 128, 0, ctx->stream>>>(off, cnt, cap, m->max_points, bucket, kept);
  VFM_TRY(launch_check(ctx, "bucket_select_kernel"));
  exclusive_scan_kernel<<<1, 1024, 0, ctx->stream>>>(kept, m->start, (int)cap, scal + 2);
  VFM_TRY(launch_check(ctx, "exclusive_scan_kernel"));
  VFM_CUDA(cudaMemcpyAsync(m->num, kept, sizeof(int32_t) * cap, cudaMemcpyDeviceToDevice, ctx->stream));
  int32_t scal_h[4];
  VFM_CUDA(cudaMemcpyAsync(scal_h, scal, sizeof(scal_h), cudaMemcpyDeviceToHost, ctx->stream));
  VFM_CUDA(cudaStreamSynchronize(ctx->stream));
  VFM_CHECK_ARG(scal_h[0] == 0, "voxel_map_build: %d point(s) outside the +-2^20 voxel range or not finite", scal_h[0]);
  m->n_points = scal_h[2];
  VFM_CUDA(cudaMalloc(&m->pts, sizeof(double) * 3 * (size_t)(m->n_points > 0 ? m->n_points : 1)));
  VFM_CUDA(cudaMalloc(&m->src_idx, sizeof(int32_t) * (size_t)(m->n_points > 0 ? m->n_points : 1)));
  map_fill_kernel<<<ceil_div(cap, 128), 128, 0, ctx->stream>>>(xyz, off, bucket, kept, m->start, cap, m->pts, m->src_idx);
  VFM_TRY(launch_check(ctx, "map_fill_kernel"));
  VFM_CUDA(cudaStreamSynchronize(ctx->stream));   // scratch (off, bucket, kept) is released with the arena
  return VFMREG_OK;
}

static MapView view_of(const vfmreg_voxel_map* m) {
  MapView v;
  v.keys = m->keys;
  v.start = m->start;
  v.num = m->num;
  v.pts = m->pts;
  v.mask = m->cap - 1;
  v.vs = m->voxel_size;
  return v;
}

int voxel_map_nearest(vfmreg_ctx* ctx, const vfmreg_voxel_map* m, const double* q, int64_t n, double max_dist, int32_t* nn, double* d2) {
  VFM_CHECK_ARG(m && m->n_points > 0, "voxel_map_nearest: empty map");
  VFM_CHECK_ARG(n > 0, "voxel_map_nearest: no query points");
  map_nearest_kernel<<<ceil_div(n * 32, 128), 128, 0, ctx->stream>>>(view_of(m), q, (int)n, max_dist, nn, d2);
  return launch_check(ctx, "map_nearest_kernel");
}

int icp_register_frame(vfmreg_ctx* ctx, const vfmreg_voxel_map* m, const double* frame, int64_t n, const double* T0, double max_dist,
                       double kernel, int max_iters, double* T_out, int32_t* iters_out, int32_t* corr_out) {
  // RegisterFrame: "if (voxel_map.Empty()) return initial_guess" (Registration.cpp:150)
  if (!m || m->n_points == 0 || n == 0) {
    memcpy(T_out, T0, 16 * sizeof(double));
    if (iters_out) *iters_out = 0;
    if (corr_out) *corr_out = 0;
    return VFMREG_OK;
  }
  const int64_t want_blocks = ceil_div(n * 32, ICP_THREADS);
  const int blocks = (int)(want_blocks < ctx->sm_count * 8 ? want_blocks : ctx->sm_count * 8);
  double* src = arena_take<double>(ctx, (size_t)n * 3);
  double* partial = arena_take<double>(ctx, (size_t)blocks * ICP_TERMS);
  IcpState* st = arena_take<IcpState>(ctx, 1);
  if (!src || !partial || !st) {
    set_error("register_frame: scratch arena too small");
    return VFMREG_ERR_ALLOC;
  }
  IcpState h;
  memset(&h, 0, sizeof(h));
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) h.est[i * 3 + j] = h.T[i * 3 + j] = T0[i * 4 + j];
    h.est[9 + i] = h.T[9 + i] = T0[i * 4 + 3];
  }
  VFM_CUDA(cudaMemcpyAsync(st, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
  VFM_CUDA(cudaMemcpyAsync(src, frame, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
  const MapView view = view_of(m);
  int issued = 0;
  while (issued < max_iters) {
    const int chunk = (max_iters - issued) < 8 ? (max_iters - issued) : 8;   // a few iterations per host round trip
    for (int it = 0; it < chunk; ++it) {
      icp_accumulate_kernel<<<blocks, ICP_THREADS, 0, ctx->stream>>>(view, src, (int)n, max_dist, kernel, st, partial);
      VFM_TRY(launch_check(ctx, "icp_accumulate_kernel"));
      icp_solve_kernel<<<1, 32, 0, ctx->stream>>>(partial, blocks, st);
      VFM_TRY(launch_check(ctx, "icp_solve_kernel"));
    }
    issued += chunk;
    VFM_CUDA(cudaMemcpyAsync(&h, st, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    VFM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h.done) break;
  }
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T_out[i * 4 + j] = h.T[i * 3 + j];
    T_out[i * 4 + 3] = h.T[9 + i];
  }
  T_out[12] = T_out[13] = T_out[14] = 0.0;
  T_out[15] = 1.0;
  if (iters_out) *iters_out = h.iters;
  if (corr_out) *corr_out = h.last_corr;
  return VFMREG_OK;
}

// VFM-ICP (Registration.cpp:197-382): the correspondence-driven loop above, then the vanilla loop with what is left of
// the iteration budget.  vfm_src / vfm_tgt: the descriptor correspondences (raw source coordinates, map coordinates).
int icp_register_frame_vfm(vfmreg_ctx* ctx, const vfmreg_voxel_map* m, const double* frame, int64_t n, const double* vfm_src,
                           const double* vfm_tgt, int64_t k, const double* T0, double max_dist, double kernel, int max_iters,
                           double* T_out, int32_t* vfm_iters, int32_t* iters_out, int32_t* vfm_kept) {
  if (!m || m->n_points == 0) {   // "if (voxel_map.EmptyN()) return initial_guess"
    memcpy(T_out, T0, 16 * sizeof(double));
    if (vfm_iters) *vfm_iters = 0;
    if (iters_out) *iters_out = 0;
    if (vfm_kept) *vfm_kept = 0;
    return VFMREG_OK;
  }
  VFM_CHECK_ARG(k >= 0 && k < (1 << 20), "register_frame_vfm: too many descriptor correspondences");
  int k2 = 1;
  while (k2 < k) k2 <<= 1;
  const size_t kk = (size_t)(k > 0 ? k : 1);
  double* bufs = arena_take<double>(ctx, kk * 3 * 4 + kk + (size_t)k2 + 16);
  Phase1Out* out = arena_take<Phase1Out>(ctx, 1);
  double* T0d = arena_take<double>(ctx, 16);
  if (!bufs || !out || !T0d) {
    set_error("register_frame_vfm: scratch arena too small");
    return VFMREG_ERR_ALLOC;
  }
  VFM_CUDA(cudaMemcpyAsync(T0d, T0, 16 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  vfm_icp_phase1_kernel<<<1, P1_THREADS, 0, ctx->stream>>>(vfm_src, vfm_tgt, (int)k, T0d, kernel, max_iters, bufs, bufs + kk * 3,
                                                          bufs + kk * 6, bufs + kk * 9, bufs + kk * 12, bufs + kk * 13, out);
  VFM_TRY(launch_check(ctx, "vfm_icp_phase1_kernel"));
  Phase1Out h;
  VFM_CUDA(cudaMemcpyAsync(&h, out, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  VFM_CUDA(cudaStreamSynchronize(ctx->stream));
  if (vfm_iters) *vfm_iters = h.iters;
  if (vfm_kept) *vfm_kept = h.kept;
  double T1[16] = {0};
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T1[i * 4 + j] = h.T[i * 3 + j];
    T1[i * 4 + 3] = h.T[9 + i];
  }
  T1[15] = 1.0;
  const int left = max_iters - h.iters;
  if (left <= 0 || n == 0) {
    memcpy(T_out, T1, sizeof(T1));
    if (iters_out) *iters_out = 0;
    return VFMREG_OK;
  }
  arena_reset(ctx);
  return icp_register_frame(ctx, m, frame, n, T1, max_dist, kernel, left, T_out, iters_out, nullptr);
}

}  // namespace vfm

using namespace vfm;

extern "C" {

int vfmreg_voxel_downsample(vfmreg_ctx* ctx, const void* points, int64_t n, int32_t cols, int32_t elem_size, double voxel_size,
                            int32_t* keep_idx, int32_t* count) {
  VFM_CHECK_ARG(ctx && points && keep_idx && count, "voxel_downsample: null pointer");
  VFM_CUDA(cudaSetDevice(ctx->device));
  arena_reset(ctx);
  VFM_TRY(arena_reserve(ctx, voxel_scratch(n)));
  return voxel_downsample(ctx, points, n, cols, elem_size, voxel_size, keep_idx, count);
}

int vfmreg_gather_rows(vfmreg_ctx* ctx, const void* rows, int32_t row_bytes, const int32_t* idx, const int32_t* count,
                       int64_t max_rows, void* out) {
  VFM_CHECK_ARG(ctx && rows && idx && count && out, "gather_rows: null pointer");
  VFM_CUDA(cudaSetDevice(ctx->device));
  return gather_rows_generic(ctx, rows, row_bytes, idx, count, max_rows, out);
}

int vfmreg_voxel_map_create(vfmreg_ctx* ctx, double voxel_size, int32_t max_points_per_voxel, vfmreg_voxel_map** out) {
  VFM_CHECK_ARG(ctx && out, "voxel_map_create: null pointer");
  VFM_CHECK_ARG(voxel_size > 0.0 && max_points_per_voxel > 0, "voxel_map_create: voxel_size and max_points_per_voxel must be > 0");
  vfmreg_voxel_map* m = new vfmreg_voxel_map();
  m->voxel_size = voxel_size;
  m->max_points = max_points_per_voxel;
  m->device = ctx->device;
  *out = m;
  return VFMREG_OK;
}

void vfmreg_voxel_map_destroy(vfmreg_voxel_map* map) {
  if (!map) return;
  cudaSetDevice(map->device);
  cudaDeviceSynchronize();
  map_free(map);
  delete map;
}

int vfmreg_voxel_map_build(vfmreg_ctx* ctx, vfmreg_voxel_map* map, const double* xyz, int64_t n) {
  VFM_CHECK_ARG(ctx && map && (xyz || n == 0), "voxel_map_build: null pointer");
  VFM_CUDA(cudaSetDevice(ctx->device));
  arena_reset(ctx);
  VFM_TRY(arena_reserve(ctx, voxel_scratch(n)));
  return voxel_map_build(ctx, map, xyz, n);
}

int64_t vfmreg_voxel_map_size(const vfmreg_voxel_map* map) { return map ? map->n_points : 0; }

int vfmreg_voxel_map_points(vfmreg_ctx* ctx, const vfmreg_voxel_map* map, double* xyz_out, int32_t* src_idx_out) {
  VFM_CHECK_ARG(ctx && map, "voxel_map_points: null pointer");
  if (map->n_points == 0) return VFMREG_OK;
  VFM_CUDA(cudaSetDevice(ctx->device));
  if (xyz_out)
    VFM_CUDA(cudaMemcpyAsync(xyz_out, map->pts, sizeof(double) * 3 * (size_t)map->n_points, cudaMemcpyDeviceToDevice, ctx->stream));
  if (src_idx_out)
    VFM_CUDA(cudaMemcpyAsync(src_idx_out, map->src_idx, sizeof(int32_t) * (size_t)map->n_points, cudaMemcpyDeviceToDevice, ctx->stream));
  return VFMREG_OK;
}

int vfmreg_voxel_map_nearest(vfmreg_ctx* ctx, const vfmreg_voxel_map* map, const double* query, int64_t n, double max_dist,
                             int32_t* nn_idx, double* nn_d2) {
  VFM_CHECK_ARG(ctx && map && query && nn_idx, "voxel_map_nearest: null pointer");
  VFM_CUDA(cudaSetDevice(ctx->device));
  return voxel_map_nearest(ctx, map, query, n, max_dist, nn_idx, nn_d2);
}

int vfmreg_register_frame(vfmreg_ctx* ctx, const vfmreg_voxel_map* map, const double* frame, int64_t n, const double* T0,
                          double max_correspondence_distance, double kernel, int32_t max_iterations, double* T_out,
                          int32_t* iterations, int32_t* correspondences) {
  VFM_CHECK_ARG(ctx && T0 && T_out && (frame || n == 0), "register_frame: null pointer");
  VFM_CHECK_ARG(max_iterations > 0, "register_frame: max_iterations must be > 0");
  VFM_CUDA(cudaSetDevice(ctx->device));
  arena_reset(ctx);
  VFM_TRY(arena_reserve(ctx, arena_bytes((size_t)(n > 0 ? n : 1) * 3, 8) + arena_bytes((size_t)(ctx->sm_count * 8 + 1) * ICP_TERMS, 8) + 4096));
  return icp_register_frame(ctx, map, frame, n, T0, max_correspondence_distance, kernel, max_iterations, T_out, iterations,
                            correspondences);
}

int vfmreg_register_frame_vfm(vfmreg_ctx* ctx, const vfmreg_voxel_map* map, const double* frame, int64_t n, const double* vfm_src,
                              const double* vfm_tgt, int64_t k, const double* T0, double max_correspondence_distance, double kernel,
                              int32_t max_iterations, double* T_out, int32_t* vfm_iterations, int32_t* iterations,
                              int32_t* vfm_kept) {
  VFM_CHECK_ARG(ctx && T0 && T_out && (frame || n == 0) && ((vfm_src && vfm_tgt) || k == 0), "register_frame_vfm: null pointer");
  VFM_CHECK_ARG(max_iterations > 0, "register_frame_vfm: max_iterations must be > 0");
  VFM_CUDA(cudaSetDevice(ctx->device));
  arena_reset(ctx);
  size_t k2 = 1;
  while ((int64_t)k2 < k) k2 <<= 1;
  const size_t need1 = arena_bytes((size_t)(k > 0 ? k : 1) * 13 + k2 + 16, 8) + 4096;
  const size_t need2 = arena_bytes((size_t)(n > 0 ? n : 1) * 3, 8) + arena_bytes((size_t)(ctx->sm_count * 8 + 1) * ICP_TERMS, 8) + 4096;
  VFM_TRY(arena_reserve(ctx, need1 > need2 ? need1 : need2));
  return icp_register_frame_vfm(ctx, map, frame, n, vfm_src, vfm_tgt, k, T0, max_correspondence_distance, kernel, max_iterations,
                                T_out, vfm_iterations, iterations, vfm_kept);
}

}  // extern "C"
