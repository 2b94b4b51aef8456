// TEASER++-style certifiable-style robust solve, the alternative solver of the reference's experiment driver
// (registration_node.py:91-131: teaserpp_python.RobustRegistrationSolver with cbar2 = 1, noise_bound = 0.2,
// estimate_scaling = False, PMC_EXACT, CHAIN, GNC_TLS, gnc_factor 1.4, 10000 iterations, cost threshold 1e-16).
// TEASER++ is not in the reference tree (cloned at HEAD by its Dockerfile:74); this follows the published algorithm
// (Yang, Shi, Carlone, "TEASER: Fast and Certifiable Point Cloud Registration", T-RO 2020, and teaser/registration.cc as
// recalled) -- "parity unpinned", like the other third-party arithmetic of the path:
//   1. translation-invariant measurements: correspondences i, j are compatible iff | ||b_i - b_j|| - ||a_i - a_j|| | <= beta,
//      beta = 2 noise_bound sqrt(cbar2) (scale fixed to 1)                       -> K x K compatibility graph (GPU, O(K^2))
//   2. maximum clique of that graph (exact branch and bound with greedy-colouring bounds on bitsets, host; node budget)
//   3. rotation: GNC-TLS on the chain of TIMs of the sorted clique (weighted Kabsch per iteration), host float64
//   4. translation: component-wise TLS by adaptive voting over the clique's b_i - R a_i, host float64
// The quadratic part runs on the GPU; 2-4 touch a few thousand values and are sequential by nature.
// THIS FILE IS COMPILED WITH -fmad=false (the graph test must not depend on contraction).
#include <math.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace vfm {

constexpr int TS_MAX_K = 16384;

// bit j of row i: correspondences i and j are compatible (i != j)
__global__ void __launch_bounds__(256)
    tim_graph_kernel(const double* __restrict__ src, const double* __restrict__ tgt, int k, int words, double beta, uint32_t* __restrict__ adj) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)k * words) return;
  const int i = (int)(g / words), w = (int)(g % words);
  const double ax = src[3 * i], ay = src[3 * i + 1], az = src[3 * i + 2];
  const double bx = tgt[3 * i], by = tgt[3 * i + 1], bz = tgt[3 * i + 2];
  uint32_t bits = 0;
  for (int b = 0; b < 32; ++b) {
    const int j = w * 32 + b;
    if (j >= k || j == i) continue;
    const double dax = src[3 * j] - ax, day = src[3 * j + 1] - ay, daz = src[3 * j + 2] - az;
    const double dbx = tgt[3 * j] - bx, dby = tgt[3 * j + 1] - by, dbz = tgt[3 * j + 2] - bz;
    const double na = sqrt((dax * dax + day * day) + daz * daz), nb = sqrt((dbx * dbx + dby * dby) + dbz * dbz);
    if (fabs(na - nb) <= beta) bits |= 1u << b;
  }
  adj[g] = bits;
}

// ---- maximum clique (host): vertices relabelled by (degree descending, index ascending); MCQ-style branch and bound:
// candidates are greedily coloured, a vertex whose colour number cannot lift the clique past the best one is cut.
struct CliqueSolver {
  int n = 0, words = 0;
  std::vector<uint64_t> adj;   // relabelled adjacency, n x words
  std::vector<int> best, cur;
  long long nodes = 0, budget = 0;
  bool exhausted = false;

  const uint64_t* row(int v) const { return adj.data() + (size_t)v * words; }

  void expand(std::vector<uint64_t>& P) {
    if (exhausted) return;
    if (++nodes > budget) {
      exhausted = true;
      return;
    }
    // greedy sequential colouring of P in index order: order[] lists the vertices colour class by colour class
    std::vector<int> order, colour;
    std::vector<uint64_t> U = P, Q(words);
    int c = 0;
    for (;;) {
      bool any = false;
      for (int w = 0; w < words && !any; ++w) any = U[w] != 0;
      if (!any) break;
      ++c;
      Q = U;
      for (int w = 0; w < words; ++w) {
        while (Q[w]) {   // clearing bits never sets lower ones: the scan only moves upwards
          const int v = w * 64 + __builtin_ctzll(Q[w]);
          order.push_back(v);
          colour.push_back(c);
          U[v >> 6] &= ~(1ull << (v & 63));
          Q[v >> 6] &= ~(1ull << (v & 63));
          const uint64_t* rv = row(v);
          for (int x = w; x < words; ++x) Q[x] &= ~rv[x];   // neighbours of v cannot share its colour
        }
      }
    }
    for (int t = (int)order.size() - 1; t >= 0; --t) {
      if ((int)cur.size() + colour[t] <= (int)best.size()) return;   // colour bound
      const int v = order[t];
      cur.push_back(v);
      std::vector<uint64_t> Pn(words);
      bool nonempty = false;
      const uint64_t* rv = row(v);
      for (int w = 0; w < words; ++w) {
        Pn[w] = P[w] & rv[w];
        nonempty |= Pn[w] != 0;
      }
      if (nonempty)
        expand(Pn);
      else if (cur.size() > best.size())
        best = cur;
      cur.pop_back();
      P[v >> 6] &= ~(1ull << (v & 63));
      if (exhausted) return;
    }
  }
};

// maximum clique of the k-vertex graph given as 32-bit adjacency rows; returns the clique in ascending original indices
static std::vector<int> max_clique(const std::vector<uint32_t>& adj32, int k, int words32, long long budget, bool* exact) {
  std::vector<int> deg(k, 0);
  for (int i = 0; i < k; ++i)
    for (int w = 0; w < words32; ++w) deg[i] += __builtin_popcount(adj32[(size_t)i * words32 + w]);
  std::vector<int> perm(k);   // new label -> original index
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return deg[a] > deg[b]; });
  std::vector<int> inv(k);
  for (int v = 0; v < k; ++v) inv[perm[v]] = v;
  CliqueSolver s;
  s.n = k;
  s.words = (k + 63) / 64;
  s.adj.assign((size_t)k * s.words, 0);
  for (int i = 0; i < k; ++i)
    for (int w = 0; w < words32; ++w) {
      uint32_t bits = adj32[(size_t)i * words32 + w];
      while (bits) {
        const int j = w * 32 + __builtin_ctz(bits);
        bits &= bits - 1;
        const int a = inv[i], b = inv[j];
        s.adj[(size_t)a * s.words + (b >> 6)] |= 1ull << (b & 63);
      }
    }
  // greedy clique in the new order as the starting bound
  for (int v = 0; v < k; ++v) {
    bool ok = true;
    for (int u : s.best) ok = ok && ((s.row(u)[v >> 6] >> (v & 63)) & 1ull);
    if (ok) s.best.push_back(v);
  }
  std::vector<uint64_t> P(s.words, 0);
  for (int v = 0; v < k; ++v) P[v >> 6] |= 1ull << (v & 63);
  s.budget = budget;
  s.expand(P);
  *exact = !s.exhausted;
  std::vector<int> out;
  for (int v : s.best) out.push_back(perm[v]);
  std::sort(out.begin(), out.end());
  return out;
}

// ---- rotation: GNC-TLS (teaser GNCTLSRotationSolver::solveForRotation) on N measurement pairs a -> b
bool host_fit_from_sigma(const double* S, const double* pm, const double* qm, double* rt);   // ransac.cu (canonical Kabsch)

static void gnc_tls_rotation(const std::vector<double>& a, const std::vector<double>& b, int n, double noise_bound, double gnc_factor,
                             double cost_threshold, int max_iterations, double R[9], std::vector<uint8_t>* inliers, int* iterations) {
  const double nb2 = noise_bound * noise_bound;
  std::vector<double> w(n, 1.0), r2(n, 0.0);
  double mu = 1.0, prev_cost = INFINITY;
  for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  const double zero[3] = {0, 0, 0};
  int it = 0;
  for (; it < max_iterations; ++it) {
    double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < n; ++j)
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) S[r * 3 + c] += w[j] * b[3 * j + r] * a[3 * j + c];
    double rt[12];
    if (host_fit_from_sigma(S, zero, zero, rt))
      for (int i = 0; i < 9; ++i) R[i] = rt[i];
    double max_r = 0.0;
    for (int j = 0; j < n; ++j) {
      double e = 0.0;
      for (int r = 0; r < 3; ++r) {
        const double d = b[3 * j + r] - ((R[r * 3] * a[3 * j] + R[r * 3 + 1] * a[3 * j + 1]) + R[r * 3 + 2] * a[3 * j + 2]);
        e += d * d;
      }
      r2[j] = e;
      max_r = e > max_r ? e : max_r;
    }
    if (it == 0) {
      mu = 1.0 / (2.0 * max_r / nb2 - 1.0);
      if (mu <= 0.0) break;   // every residual is already inside the bound
    }
    const double th1 = (mu + 1.0) / mu * nb2, th2 = mu / (mu + 1.0) * nb2;
    double cost = 0.0;
    for (int j = 0; j < n; ++j) {
      cost += w[j] * r2[j];
      if (r2[j] >= th1)
        w[j] = 0.0;
      else if (r2[j] <= th2)
        w[j] = 1.0;
      else
        w[j] = sqrt(nb2 * mu * (mu + 1.0) / r2[j]) - mu;
    }
    const double diff = fabs(cost - prev_cost);
    mu *= gnc_factor;
    prev_cost = cost;
    if (diff < cost_threshold) {
      ++it;
      break;
    }
  }
  *iterations = it;
  inliers->assign(n, 0);
  for (int j = 0; j < n; ++j) (*inliers)[j] = w[j] >= 0.5;
}

// ---- translation: scalar TLS by adaptive voting (teaser ScalarTLSEstimator::estimate), ranges all equal to `range`
static double tls_scalar(const std::vector<double>& x, double range, int* n_inliers) {
  const int n = (int)x.size();
  std::vector<std::pair<double, int>> ev(2 * n);   // (value, +(i+1) entering | -(i+1) leaving)
  for (int i = 0; i < n; ++i) {
    ev[2 * i] = {x[i] - range, i + 1};
    ev[2 * i + 1] = {x[i] + range, -(i + 1)};
  }
  std::stable_sort(ev.begin(), ev.end(), [](const std::pair<double, int>& p, const std::pair<double, int>& q) { return p.first < q.first; });
  const double wgt = 1.0 / (range * range);
  double ranges_inverse_sum = range * n, dot_xw = 0.0, dot_w = 0.0, sum_x = 0.0, sum_x2 = 0.0, best_cost = INFINITY, best = 0.0;
  int card = 0;
  for (int e = 0; e < 2 * n; ++e) {
    const int idx = abs(ev[e].second) - 1;
    const double eps = ev[e].second > 0 ? 1.0 : -1.0;
    card += ev[e].second > 0 ? 1 : -1;
    dot_w += eps * wgt;
    dot_xw += eps * wgt * x[idx];
    ranges_inverse_sum -= eps * range;
    sum_x += eps * x[idx];
    sum_x2 += eps * x[idx] * x[idx];
    if (card <= 0) continue;
    const double xh = dot_xw / dot_w;
    const double cost = (card * xh * xh + sum_x2 - 2.0 * sum_x * xh) + ranges_inverse_sum;
    if (cost < best_cost) {
      best_cost = cost;
      best = xh;
    }
  }
  int cnt = 0;
  for (int i = 0; i < n; ++i) cnt += fabs(x[i] - best) <= range;
  *n_inliers = cnt;
  return best;
}

}  // namespace vfm

using namespace vfm;

extern "C" int vfmreg_teaser_solve(vfmreg_ctx* ctx, const double* src_xyz, const double* tgt_xyz, int64_t k, const vfmreg_teaser_params* p,
                                   double* T, int32_t* clique, int32_t* stats) {
  VFM_CHECK_ARG(ctx && src_xyz && tgt_xyz && p && T, "teaser_solve: null pointer");
  VFM_CHECK_ARG(k >= 0 && k <= TS_MAX_K, "teaser_solve: between 0 and %d correspondences supported, got %lld", TS_MAX_K, (long long)k);
  VFM_CHECK_ARG(p->noise_bound > 0 && p->cbar2 > 0 && p->gnc_factor > 1 && p->max_iterations > 0, "teaser_solve: bad parameters");
  for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
  if (stats) stats[0] = stats[1] = stats[2] = stats[3] = stats[4] = 0;
  if (k < 2) return VFMREG_OK;   // TEASER: a clique of one vertex is no solution; identity
  VFM_CUDA(cudaSetDevice(ctx->device));
  const int kk = (int)k, words = (kk + 31) / 32;
  arena_reset(ctx);
  VFM_TRY(arena_reserve(ctx, 2 * arena_bytes((size_t)kk * 3, 8) + arena_bytes((size_t)kk * words, 4) + 4096));
  double* d_src = arena_take<double>(ctx, (size_t)kk * 3);
  double* d_tgt = arena_take<double>(ctx, (size_t)kk * 3);
  uint32_t* d_adj = arena_take<uint32_t>(ctx, (size_t)kk * words);
  VFM_CHECK_ARG(d_src && d_tgt && d_adj, "teaser_solve: scratch arena too small");
  VFM_CUDA(cudaMemcpyAsync(d_src, src_xyz, (size_t)kk * 24, cudaMemcpyHostToDevice, ctx->stream));
  VFM_CUDA(cudaMemcpyAsync(d_tgt, tgt_xyz, (size_t)kk * 24, cudaMemcpyHostToDevice, ctx->stream));
  const double beta = 2.0 * p->noise_bound * sqrt(p->cbar2);
  tim_graph_kernel<<<ceil_div((long long)kk * words, 256), 256, 0, ctx->stream>>>(d_src, d_tgt, kk, words, beta, d_adj);
  VFM_TRY(launch_check(ctx, "tim_graph_kernel"));
  std::vector<uint32_t> adj((size_t)kk * words);
  VFM_CUDA(cudaMemcpyAsync(adj.data(), d_adj, adj.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
  VFM_CUDA(cudaStreamSynchronize(ctx->stream));
  bool exact = true;
  const std::vector<int> cl = max_clique(adj, kk, words, p->max_clique_nodes > 0 ? p->max_clique_nodes : 20000000LL, &exact);
  const int n = (int)cl.size();
  if (stats) {
    stats[0] = n;
    stats[1] = exact ? 1 : 0;
  }
  if (clique)
    for (int i = 0; i < n; ++i) clique[i] = cl[i];
  if (n <= 1) return VFMREG_OK;   // "max clique size <= 1: abort" (identity)
  // CHAIN: TIMs of consecutive clique members
  std::vector<double> a(3 * (n - 1)), b(3 * (n - 1));
  for (int i = 0; i + 1 < n; ++i)
    for (int c = 0; c < 3; ++c) {
      a[3 * i + c] = src_xyz[3 * cl[i + 1] + c] - src_xyz[3 * cl[i] + c];
      b[3 * i + c] = tgt_xyz[3 * cl[i + 1] + c] - tgt_xyz[3 * cl[i] + c];
    }
  double R[9];
  std::vector<uint8_t> rot_in;
  int iters = 0;
  gnc_tls_rotation(a, b, n - 1, p->noise_bound, p->gnc_factor, p->cost_threshold, p->max_iterations, R, &rot_in, &iters);
  double t[3];
  int trans_in = n;
  for (int c = 0; c < 3; ++c) {
    std::vector<double> x(n);
    for (int i = 0; i < n; ++i) {
      const double* s = src_xyz + 3 * cl[i];
      x[i] = tgt_xyz[3 * cl[i] + c] - ((R[c * 3] * s[0] + R[c * 3 + 1] * s[1]) + R[c * 3 + 2] * s[2]);
    }
    int cnt = 0;
    t[c] = tls_scalar(x, p->noise_bound * sqrt(p->cbar2), &cnt);
    trans_in = cnt < trans_in ? cnt : trans_in;
  }
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T[r * 4 + c] = R[r * 3 + c];
    T[r * 4 + 3] = t[r];
  }
  if (stats) {
    stats[2] = iters;
    stats[3] = (int)std::count(rot_in.begin(), rot_in.end(), (uint8_t)1);
    stats[4] = trans_in;
  }
  return VFMREG_OK;
}
