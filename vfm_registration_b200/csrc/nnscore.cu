// SURVEY 8f row 4 -- the hypothesis score Open3D 0.18 actually uses inside registration_ransac_based_on_correspondence
// (reference call site registration_node.py:312-327; control flow as recalled in SURVEY A.8, "parity unpinned"): every
// hypothesis transforms the WHOLE source cloud and is scored by the nearest target point of every transformed point --
//   fitness = #(points whose nearest target point is closer than max_dist) / #points,  rmse = sqrt(sum d^2 / #those),
//   best = larger fitness, then smaller rmse, then lower hypothesis id; the winning 3-point transform is returned as is.
// With the reference's literal max_dist = 10000 every point counts, so the winner is the hypothesis with the smallest
// scan -> map chamfer RMSE.
//
// The nearest neighbour must be exact at any distance (a wrong hypothesis throws the scan far from the map), so the
// 27-voxel search of voxel.cu does not do: the target cloud gets a balanced k-d tree (built once on the host: median
// splits along the widest axis, leaves of <= 16 points, implicit heap order) that stays in L2 (24 B per point), and every
// query walks it with a short stack.  float64 throughout, THIS FILE IS COMPILED WITH -fmad=false.
//
// Kernels
//   kdtree_nearest_kernel  one thread per query (the standalone nearest-neighbour entry point)
//   nn_score_kernel        one CTA per hypothesis; thread t takes source points t, t + 128, ...; count and sum d^2 are
//                          reduced in a fixed order (deterministic).  Bound: L2 latency of the tree walk.
//   nn_best_kernel         arg-best over the hypotheses, one CTA
#include <algorithm>
#include <vector>

#include "common.cuh"

struct vfmreg_kdtree {
  vfmreg_ctx* ctx = nullptr;
  int64_t n = 0;
  int depth = 0;              // leaves sit at this depth: 2^depth leaves, 2^depth - 1 internal nodes
  double* pts = nullptr;      // (n, 3) points in leaf order
  int32_t* perm = nullptr;    // leaf order -> caller's index
  int32_t* leaf_start = nullptr;   // 2^depth + 1 offsets into pts
  double* split_val = nullptr;     // per internal node (heap order)
  int8_t* split_dim = nullptr;
};

namespace vfm {

constexpr int KD_LEAF = 16;
constexpr int KD_MAX_DEPTH = 26;
constexpr int NN_THREADS = 128;

struct KdView {
  const double* pts;
  const int32_t* leaf_start;
  const double* split_val;
  const int8_t* split_dim;
  int depth;
};

// exact nearest neighbour of q among the tree's points with d^2 < limit2 (strict); returns the leaf-order position or -1
__device__ __forceinline__ int kd_nearest(const KdView& t, double qx, double qy, double qz, double limit2, double& best2) {
  int best = -1;
  best2 = limit2;
  int st_node[KD_MAX_DEPTH + 2];
  double st_bound[KD_MAX_DEPTH + 2];
  int sp = 0;
  st_node[0] = 0;
  st_bound[0] = 0.0;
  sp = 1;
  const int first_leaf = (1 << t.depth) - 1;
  while (sp > 0) {
    --sp;
    int node = st_node[sp];
    if (st_bound[sp] >= best2) continue;
    while (node < first_leaf) {   // descend towards the query, remember the other side with its plane distance
      const int dim = t.split_dim[node];
      const double diff = (dim == 0 ? qx : (dim == 1 ? qy : qz)) - t.split_val[node];
      const int near = 2 * node + (diff < 0.0 ? 1 : 2), far = 2 * node + (diff < 0.0 ? 2 : 1);
      const double b = diff * diff;
      if (b < best2) {
        st_node[sp] = far;
        st_bound[sp] = b;
        ++sp;
      }
      node = near;
    }
    const int leaf = node - first_leaf;
    const int p0 = t.leaf_start[leaf], p1 = t.leaf_start[leaf + 1];
    for (int p = p0; p < p1; ++p) {
      const double dx = t.pts[3 * p] - qx, dy = t.pts[3 * p + 1] - qy, dz = t.pts[3 * p + 2] - qz;
      const double d2 = (dx * dx + dy * dy) + dz * dz;
      if (d2 < best2) {
        best2 = d2;
        best = p;
      }
    }
  }
  return best;
}

__global__ void kdtree_nearest_kernel(KdView t, const int32_t* __restrict__ perm, const double* __restrict__ q, int64_t n, double limit2,
                                      int32_t* __restrict__ idx, double* __restrict__ d2_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double d2;
  const int p = kd_nearest(t, q[3 * i], q[3 * i + 1], q[3 * i + 2], limit2, d2);
  idx[i] = p >= 0 ? perm[p] : -1;
  if (d2_out) d2_out[i] = p >= 0 ? d2 : -1.0;
}

// one CTA per hypothesis: fitness count and sum of squared nearest-neighbour distances over the whole source cloud
__global__ void __launch_bounds__(NN_THREADS)
    nn_score_kernel(KdView t, const double* __restrict__ src, int n_src, const double* __restrict__ rts, const int32_t* __restrict__ counts_in,
                    double limit2, int32_t* __restrict__ cnt_out, double* __restrict__ sum_out) {
  __shared__ double s_sum[NN_THREADS];
  __shared__ int s_cnt[NN_THREADS];
  const int h = blockIdx.x;
  if (counts_in[h] < 0) {   // degenerate sample: never a candidate
    if (threadIdx.x == 0) {
      cnt_out[h] = -1;
      sum_out[h] = 0.0;
    }
    return;
  }
  double rt[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) rt[i] = rts[(int64_t)h * 12 + i];
  int cnt = 0;
  double sum = 0.0;
  for (int i = threadIdx.x; i < n_src; i += NN_THREADS) {
    const double px = src[3 * i], py = src[3 * i + 1], pz = src[3 * i + 2];
    const double x = fma(rt[0], px, fma(rt[1], py, fma(rt[2], pz, rt[9])));
    const double y = fma(rt[3], px, fma(rt[4], py, fma(rt[5], pz, rt[10])));
    const double z = fma(rt[6], px, fma(rt[7], py, fma(rt[8], pz, rt[11])));
    double d2;
    if (kd_nearest(t, x, y, z, limit2, d2) >= 0) {
      ++cnt;
      sum += d2;
    }
  }
  s_sum[threadIdx.x] = sum;
  s_cnt[threadIdx.x] = cnt;
  __syncthreads();
  for (int off = NN_THREADS / 2; off >= 1; off >>= 1) {   // fixed tree: the same sum on every run
    if ((int)threadIdx.x < off) {
      s_sum[threadIdx.x] += s_sum[threadIdx.x + off];
      s_cnt[threadIdx.x] += s_cnt[threadIdx.x + off];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    cnt_out[h] = s_cnt[0];
    sum_out[h] = s_sum[0];
  }
}

// best = more inliers (fitness), then smaller rmse = sqrt(sum / count), then lower id; stats = {best or -1, its count}
__global__ void __launch_bounds__(256)
    nn_best_kernel(const int32_t* __restrict__ cnt, const double* __restrict__ sum, const double* __restrict__ rts, int n_hyp,
                   double* __restrict__ T, int64_t* __restrict__ stats, double* __restrict__ best_sum) {
  __shared__ int s_c[256], s_i[256];
  __shared__ double s_r[256];
  int bc = -1, bi = 0x7fffffff;
  double br = 0.0;
  for (int h = threadIdx.x; h < n_hyp; h += 256) {
    const int c = cnt[h];
    if (c <= 0) continue;
    const double r = sum[h] / (double)c;   // mean squared distance: same order as the rmse
    if (c > bc || (c == bc && (r < br || (r == br && h < bi)))) {
      bc = c;
      br = r;
      bi = h;
    }
  }
  s_c[threadIdx.x] = bc;
  s_r[threadIdx.x] = br;
  s_i[threadIdx.x] = bi;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int t = 1; t < 256; ++t) {
      const int c = s_c[t], i = s_i[t];
      const double r = s_r[t];
      if (c > bc || (c == bc && c > 0 && (r < br || (r == br && i < bi)))) {
        bc = c;
        br = r;
        bi = i;
      }
    }
    const bool ok = bc > 0;
    stats[0] = ok ? bi : -1;
    stats[1] = ok ? bc : 0;
    *best_sum = ok ? sum[bi] : 0.0;
    for (int k = 0; k < 16; ++k) {
      const int i = k >> 2, j = k & 3;
      T[k] = ok ? ((i == 3) ? ((j == 3) ? 1.0 : 0.0) : ((j == 3) ? rts[(int64_t)bi * 12 + 9 + i] : rts[(int64_t)bi * 12 + i * 3 + j]))
                : ((i == j) ? 1.0 : 0.0);
    }
  }
}

static KdView view_of(const vfmreg_kdtree* t) {
  KdView v;
  v.pts = t->pts;
  v.leaf_start = t->leaf_start;
  v.split_val = t->split_val;
  v.split_dim = t->split_dim;
  v.depth = t->depth;
  return v;
}

// hypotheses from sampled correspondences (ransac.cu)
int ransac_hypotheses(vfmreg_ctx* ctx, const void* src_xyz, const void* tgt_xyz, int xyz_f64, const int32_t* corr, const int32_t* count,
                      int32_t max_corr, const int32_t* sample_idx, int32_t n_hyp, uint64_t seed, double* rts, int32_t* counts);

}  // namespace vfm

using namespace vfm;

extern "C" {

int vfmreg_kdtree_create(vfmreg_ctx* ctx, const double* xyz_host, int64_t n, vfmreg_kdtree** out) {
  VFM_CHECK_ARG(ctx && xyz_host && out, "kdtree_create: null pointer");
  *out = nullptr;
  VFM_CHECK_ARG(n > 0 && n < (1LL << 30), "kdtree_create: bad size %lld", (long long)n);
  VFM_CUDA(cudaSetDevice(ctx->device));
  int depth = 0;
  while (((n + (1LL << depth) - 1) >> depth) > KD_LEAF && depth < KD_MAX_DEPTH) ++depth;
  const int64_t n_leaf = 1LL << depth, n_int = n_leaf - 1;
  std::vector<int32_t> order(n);
  for (int64_t i = 0; i < n; ++i) order[i] = (int32_t)i;
  std::vector<double> split_val(n_int > 0 ? n_int : 1, 0.0);
  std::vector<int8_t> split_dim(n_int > 0 ? n_int : 1, 0);
  std::vector<int32_t> leaf_start(n_leaf + 1, 0);
  // iterative build over the heap: node -> [lo, hi)
  std::vector<int64_t> lo(2 * n_leaf), hi(2 * n_leaf);
  lo[0] = 0;
  hi[0] = n;
  for (int64_t node = 0; node < n_int; ++node) {
    const int64_t a = lo[node], b = hi[node];
    int dim = 0;
    double val = 0.0;
    int64_t mid = a + (b - a + 1) / 2;
    if (b - a >= 2) {
      double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
      for (int64_t i = a; i < b; ++i)
        for (int c = 0; c < 3; ++c) {
          const double v = xyz_host[3 * (int64_t)order[i] + c];
          mn[c] = v < mn[c] ? v : mn[c];
          mx[c] = v > mx[c] ? v : mx[c];
        }
      for (int c = 1; c < 3; ++c)
        if (mx[c] - mn[c] > mx[dim] - mn[dim]) dim = c;
      std::nth_element(order.begin() + a, order.begin() + mid, order.begin() + b, [&](int32_t x, int32_t y) {
        const double vx = xyz_host[3 * (int64_t)x + dim], vy = xyz_host[3 * (int64_t)y + dim];
        return vx < vy || (vx == vy && x < y);
      });
      val = xyz_host[3 * (int64_t)order[mid] + dim];   // left: coordinates <= val, right: >= val
    } else {
      mid = b;   // 0 or 1 point: everything goes left, the right child stays empty
      val = 1e300;
    }
    split_dim[node] = (int8_t)dim;
    split_val[node] = val;
    lo[2 * node + 1] = a;
    hi[2 * node + 1] = mid;
    lo[2 * node + 2] = mid;
    hi[2 * node + 2] = b;
  }
  for (int64_t l = 0; l < n_leaf; ++l) leaf_start[l] = (int32_t)lo[n_int + l];
  leaf_start[n_leaf] = (int32_t)n;
  std::vector<double> pts(3 * n);
  for (int64_t i = 0; i < n; ++i)
    for (int c = 0; c < 3; ++c) pts[3 * i + c] = xyz_host[3 * (int64_t)order[i] + c];
  vfmreg_kdtree* t = new vfmreg_kdtree();
  t->ctx = ctx;
  t->n = n;
  t->depth = depth;
  bool ok = cudaMalloc(&t->pts, sizeof(double) * 3 * n) == cudaSuccess && cudaMalloc(&t->perm, sizeof(int32_t) * n) == cudaSuccess &&
            cudaMalloc(&t->leaf_start, sizeof(int32_t) * (n_leaf + 1)) == cudaSuccess &&
            cudaMalloc(&t->split_val, sizeof(double) * split_val.size()) == cudaSuccess &&
            cudaMalloc(&t->split_dim, split_dim.size()) == cudaSuccess;
  ok = ok && cudaMemcpy(t->pts, pts.data(), sizeof(double) * 3 * n, cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(t->perm, order.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(t->leaf_start, leaf_start.data(), sizeof(int32_t) * (n_leaf + 1), cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(t->split_val, split_val.data(), sizeof(double) * split_val.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(t->split_dim, split_dim.data(), split_dim.size(), cudaMemcpyHostToDevice) == cudaSuccess;
  if (!ok) {
    set_error("kdtree_create: device allocation / copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(t->pts); cudaFree(t->perm); cudaFree(t->leaf_start); cudaFree(t->split_val); cudaFree(t->split_dim);
    delete t;
    return VFMREG_ERR_ALLOC;
  }
  *out = t;
  return VFMREG_OK;
}

void vfmreg_kdtree_destroy(vfmreg_kdtree* t) {
  if (!t) return;
  cudaSetDevice(t->ctx->device);
  cudaDeviceSynchronize();
  cudaFree(t->pts); cudaFree(t->perm); cudaFree(t->leaf_start); cudaFree(t->split_val); cudaFree(t->split_dim);
  delete t;
}

int vfmreg_kdtree_nearest(vfmreg_ctx* ctx, const vfmreg_kdtree* tree, const double* queries, int64_t n, double max_dist, int32_t* nn_idx,
                          double* nn_d2) {
  VFM_CHECK_ARG(ctx && tree && queries && nn_idx, "kdtree_nearest: null pointer");
  VFM_CHECK_ARG(tree->ctx == ctx, "kdtree_nearest: the tree belongs to another context");
  VFM_CHECK_ARG(n >= 0 && max_dist > 0, "kdtree_nearest: bad arguments");
  if (n == 0) return VFMREG_OK;
  VFM_CUDA(cudaSetDevice(ctx->device));
  kdtree_nearest_kernel<<<ceil_div(n, 128), 128, 0, ctx->stream>>>(view_of(tree), tree->perm, queries, n, max_dist * max_dist, nn_idx, nn_d2);
  return launch_check(ctx, "kdtree_nearest_kernel");
}

int vfmreg_ransac_nn_all(vfmreg_ctx* ctx, const vfmreg_kdtree* tree, const double* src_all, int64_t n_src, const void* src_xyz,
                         const void* tgt_xyz, int xyz_f64, const int32_t* corr, const int32_t* count, int32_t max_corr,
                         const int32_t* sample_idx, int32_t n_hyp, uint64_t seed, double max_dist, double* T, int32_t* inliers,
                         double* sum_d2, int64_t* stats) {
  VFM_CHECK_ARG(ctx && tree && src_all && src_xyz && tgt_xyz && corr && count && T && stats, "ransac_nn_all: null pointer");
  VFM_CHECK_ARG(tree->ctx == ctx, "ransac_nn_all: the tree belongs to another context");
  VFM_CHECK_ARG(n_src > 0 && n_src < (1LL << 31) && n_hyp > 0 && max_corr >= 0, "ransac_nn_all: bad sizes");
  VFM_CHECK_ARG(max_dist > 0.0, "ransac_nn_all: max_dist must be > 0 (Open3D returns an empty result otherwise)");
  VFM_CUDA(cudaSetDevice(ctx->device));
  arena_reset(ctx);
  VFM_TRY(arena_reserve(ctx, arena_bytes((size_t)n_hyp * 12, 8) + arena_bytes(n_hyp, 4) * 2 + arena_bytes(n_hyp, 8) + 4096));
  double* rts = arena_take<double>(ctx, (size_t)n_hyp * 12);
  int32_t* valid = arena_take<int32_t>(ctx, n_hyp);
  int32_t* cnt = inliers ? inliers : arena_take<int32_t>(ctx, n_hyp);
  double* sum = sum_d2 ? sum_d2 : arena_take<double>(ctx, n_hyp);
  double* best_sum = arena_take<double>(ctx, 1);
  if (!rts || !valid || !cnt || !sum || !best_sum) {
    set_error("ransac_nn_all: scratch arena too small");
    return VFMREG_ERR_ALLOC;
  }
  VFM_TRY(ransac_hypotheses(ctx, src_xyz, tgt_xyz, xyz_f64, corr, count, max_corr, sample_idx, n_hyp, seed, rts, valid));
  nn_score_kernel<<<n_hyp, NN_THREADS, 0, ctx->stream>>>(view_of(tree), src_all, (int)n_src, rts, valid, max_dist * max_dist, cnt, sum);
  VFM_TRY(launch_check(ctx, "nn_score_kernel"));
  nn_best_kernel<<<1, 256, 0, ctx->stream>>>(cnt, sum, rts, n_hyp, T, stats, best_sum);
  VFM_TRY(launch_check(ctx, "nn_best_kernel"));
  // stats[2] = bit pattern of the winner's sum of squared distances (read it back as a double)
  VFM_CUDA(cudaMemcpyAsync(stats + 2, best_sum, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  return VFMREG_OK;
}

}  // extern "C"
