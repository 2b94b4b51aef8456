// Point -> pixel projection + feature gather + first-camera-wins scatter in one pass.
// Replaces project_pcl_to_image (reference dataloader/nclt.py:311-366, dataloader/oxford_robotcar.py:330-363) and the
// gather / dedup / scatter of create_descriptors (prepare_scenes.py:57-104).  The full-resolution feature map the
// reference materialises (F.interpolate to (H, W, C), image_features.py:104-108: ~0.9 GB per NCLT image) is never built:
// the value that map would hold at the integer pixel is computed from the 4 surrounding tokens.
//
// Small clouds (one kernel): one warp per point.  Lane c evaluates the projection (float64, as NumPy does) into camera c, a
// ballot picks the first camera that sees the point (first camera wins, prepare_scenes.py:97-101); the lanes then stride the
// channel dimension with 128-bit loads of the L2-resident token grid and 128-bit streaming stores of the descriptor row.
// That kernel is L2-read bound (four token rows = 4 x 4 d bytes read per 4 d bytes written: 0.24 of the HBM peak at 2 M points).
//
// Large clouds (>= PG_BINNED_MIN points): the points are binned by (camera, token cell) first --
//   classify_points_kernel  ONE THREAD per point: projection, winner camera, pixel, bilinear cell -> bin id, histogram (the
//                           warp-per-point kernel pays a warp instruction per fp64 operation of <= 16 lanes: 1.8 ms of fp64 at
//                           2 M points)
//   bin_offsets_kernel      exclusive scan over the bins (a few thousand)
//   bin_scatter_kernel      point index -> its place in the bin order
//   gather_binned_kernel    a warp takes 32 consecutive entries of the bin order; the four token rows of a cell stay in
//                           registers while consecutive entries share the cell, so every row is read from L2 about once per
//                           32 points instead of once per point; what is left is the descriptor write:
// HBM-bound: 12 B read + 4 d B written per point (unseen points are written as zeros, prepare_scenes.py:102-104).
// Both paths run the same float64 projection and the same float32 blend, operation for operation.
#include <stdlib.h>

#include "common.cuh"

namespace vfm {

constexpr int MAX_CAMS = 16;

struct CamDev {
  vfmreg_camera c;
  int64_t tok_off;
  int64_t img_off;
  int32_t bin_base;   // first (camera, token cell) bin of this camera (binned path)
  int32_t n_bins;     // total number of bins (the same in every entry)
};

struct CamPack {
  CamDev cam[MAX_CAMS];
};

__device__ __forceinline__ void axis_coeff(int dst, int n_in, int n_out, int& i0, int& i1, float& w1) {
  // PyTorch upsample_bilinear2d, align_corners=False: src = max(scale * (dst + 0.5) - 0.5, 0) in float32
  const float scale = (float)n_in / (float)n_out;
  float src = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  if (src < 0.f) src = 0.f;
  i0 = min((int)src, n_in - 1);
  i1 = i0 + ((i0 < n_in - 1) ? 1 : 0);
  w1 = __fsub_rn(src, (float)i0);
}

constexpr int PG_HIST_COPIES = 64;

__global__ void __launch_bounds__(256)
    project_gather_kernel(const float* __restrict__ points, int64_t n, CamPack pack, int n_cam, const float* __restrict__ tokens,
                          const uint8_t* __restrict__ images, int d, float* __restrict__ desc, int32_t* __restrict__ cam_of_point,
                          int32_t* __restrict__ uv) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const double x = (double)points[i * 3 + 0], y = (double)points[i * 3 + 1], z = (double)points[i * 3 + 2];
  // lane c projects the point into camera c (all cameras in parallel); the first camera that sees it wins
  bool ok = false, black = false;
  int u = 0, v = 0;
  if (lane < n_cam) {
    const vfmreg_camera& cam = pack.cam[lane].c;
    // q = P [x y z 1]^T, accumulated left to right like a NumPy row-times-column product
    const double q0 = ((cam.P[0] * x + cam.P[1] * y) + cam.P[2] * z) + cam.P[3];
    const double q1 = ((cam.P[4] * x + cam.P[5] * y) + cam.P[6] * z) + cam.P[7];
    const double q2 = ((cam.P[8] * x + cam.P[9] * y) + cam.P[10] * z) + cam.P[11];
    ok = cam.z_inclusive ? (q2 >= 0.0) : (q2 > 0.0);
    const double xf = q0 / q2 / cam.subsample;
    const double yf = q1 / q2 / cam.subsample;
    ok = ok && (fabs(xf) < 1073741824.0) && (fabs(yf) < 1073741824.0);  // also drops NaN / inf (z == 0)
    if (ok && cam.float_bounds) ok = !(xf < 0.0 || xf > (double)cam.crop_w || yf < 0.0 || yf > (double)cam.crop_h);
    const int xi = ok ? (int)xf : 0, yi = ok ? (int)yf : 0;  // truncation toward zero == ndarray.astype(int)
    ok = ok && !(xi < cam.crop_x0 || xi >= cam.crop_x0 + cam.crop_w || yi < cam.crop_y0 || yi >= cam.crop_y0 + cam.crop_h);
    u = xi - cam.crop_x0;
    v = yi - cam.crop_y0;
    if (ok && cam.black_mode && images) {
      // (u, v) index the frame the projection lives in; the stored image is un-rotated
      const int r = cam.rot90 ? u : v;
      const int cc = cam.rot90 ? (cam.img_h - 1 - v) : u;
      const int stored_w = cam.rot90 ? cam.img_h : cam.img_w;
      const uint8_t* px = images + pack.cam[lane].img_off + ((int64_t)r * stored_w + cc) * 3;
      black = (px[0] | px[1] | px[2]) == 0;
    }
    if (black && cam.black_mode == 1) ok = false;   // a black pixel hides the point from this camera
  }
  const unsigned seen = __ballot_sync(0xffffffffu, ok);
  const int found = seen ? (__ffs(seen) - 1) : -1;
  const int src_lane = found < 0 ? 0 : found;
  const int fu = __shfl_sync(0xffffffffu, u, src_lane), fv = __shfl_sync(0xffffffffu, v, src_lane);
  const bool zero_feat = __shfl_sync(0xffffffffu, (int)black, src_lane) != 0;
  if (lane == 0) {
    if (cam_of_point) cam_of_point[i] = found;
    if (uv) {
      uv[2 * i] = found >= 0 ? fu : -1;
      uv[2 * i + 1] = found >= 0 ? fv : -1;
    }
  }
  float* out = desc + i * (int64_t)d;
  const bool vec = (d & 3) == 0;
  if (found < 0 || zero_feat) {
    if (vec) {
      for (int k = 4 * lane; k < d; k += 128) __stcs(reinterpret_cast<float4*>(out + k), make_float4(0.f, 0.f, 0.f, 0.f));
    } else {
      for (int k = lane; k < d; k += 32) out[k] = 0.f;
    }
    return;
  }
  const vfmreg_camera& cam = pack.cam[found].c;
  // stored (un-rotated) pixel and stored map size
  const int pr = cam.rot90 ? fu : fv;
  const int pc = cam.rot90 ? (cam.img_h - 1 - fv) : fu;
  const int sh = cam.rot90 ? cam.img_w : cam.img_h;
  const int sw = cam.rot90 ? cam.img_h : cam.img_w;
  int y0, y1, x0, x1;
  float wy, wx;
  axis_coeff(pr, cam.grid_h, sh, y0, y1, wy);
  axis_coeff(pc, cam.grid_w, sw, x0, x1, wx);
  const float w0y = __fsub_rn(1.f, wy), w0x = __fsub_rn(1.f, wx);
  const float* g = tokens + pack.cam[found].tok_off;
  const float* t00 = g + ((int64_t)y0 * cam.grid_w + x0) * d;
  const float* t01 = g + ((int64_t)y0 * cam.grid_w + x1) * d;
  const float* t10 = g + ((int64_t)y1 * cam.grid_w + x0) * d;
  const float* t11 = g + ((int64_t)y1 * cam.grid_w + x1) * d;
#define VFM_BLEND(a, b, c_, d_) \
  __fadd_rn(__fmul_rn(w0y, __fadd_rn(__fmul_rn(w0x, a), __fmul_rn(wx, b))), __fmul_rn(wy, __fadd_rn(__fmul_rn(w0x, c_), __fmul_rn(wx, d_))))
  if (vec) {
    for (int k = 4 * lane; k < d; k += 128) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(t00 + k));
      const float4 b = __ldg(reinterpret_cast<const float4*>(t01 + k));
      const float4 c = __ldg(reinterpret_cast<const float4*>(t10 + k));
      const float4 e = __ldg(reinterpret_cast<const float4*>(t11 + k));
      float4 o;
      o.x = VFM_BLEND(a.x, b.x, c.x, e.x);
      o.y = VFM_BLEND(a.y, b.y, c.y, e.y);
      o.z = VFM_BLEND(a.z, b.z, c.z, e.z);
      o.w = VFM_BLEND(a.w, b.w, c.w, e.w);
      __stcs(reinterpret_cast<float4*>(out + k), o);
    }
  } else {
    for (int k = lane; k < d; k += 32) out[k] = VFM_BLEND(__ldg(t00 + k), __ldg(t01 + k), __ldg(t10 + k), __ldg(t11 + k));
  }
#undef VFM_BLEND
}


// ---- binned path ---------------------------------------------------------------------------------------------------
constexpr int64_t PG_BINNED_MIN = 65536;

// stored (un-rotated) pixel of (fu, fv) and its bilinear cell in the token grid
__device__ __forceinline__ void cell_of(const vfmreg_camera& cam, int fu, int fv, int& y0, int& y1, int& x0, int& x1, float& wy, float& wx) {
  const int pr = cam.rot90 ? fu : fv;
  const int pc = cam.rot90 ? (cam.img_h - 1 - fv) : fu;
  const int sh = cam.rot90 ? cam.img_w : cam.img_h;
  const int sw = cam.rot90 ? cam.img_h : cam.img_w;
  axis_coeff(pr, cam.grid_h, sh, y0, y1, wy);
  axis_coeff(pc, cam.grid_w, sw, x0, x1, wx);
}

// The projection of one point into one camera: the per-lane code of project_gather_kernel, operation for operation.
__device__ __forceinline__ bool project_one(const vfmreg_camera& cam, const uint8_t* images, int64_t img_off, double x, double y, double z,
                                            int& u, int& v, bool& black) {
  const double q0 = ((cam.P[0] * x + cam.P[1] * y) + cam.P[2] * z) + cam.P[3];
  const double q1 = ((cam.P[4] * x + cam.P[5] * y) + cam.P[6] * z) + cam.P[7];
  const double q2 = ((cam.P[8] * x + cam.P[9] * y) + cam.P[10] * z) + cam.P[11];
  bool ok = cam.z_inclusive ? (q2 >= 0.0) : (q2 > 0.0);
  const double xf = q0 / q2 / cam.subsample;
  const double yf = q1 / q2 / cam.subsample;
  ok = ok && (fabs(xf) < 1073741824.0) && (fabs(yf) < 1073741824.0);
  if (ok && cam.float_bounds) ok = !(xf < 0.0 || xf > (double)cam.crop_w || yf < 0.0 || yf > (double)cam.crop_h);
  const int xi = ok ? (int)xf : 0, yi = ok ? (int)yf : 0;
  ok = ok && !(xi < cam.crop_x0 || xi >= cam.crop_x0 + cam.crop_w || yi < cam.crop_y0 || yi >= cam.crop_y0 + cam.crop_h);
  u = xi - cam.crop_x0;
  v = yi - cam.crop_y0;
  black = false;
  if (ok && cam.black_mode && images) {
    const int r = cam.rot90 ? u : v;
    const int cc = cam.rot90 ? (cam.img_h - 1 - v) : u;
    const int stored_w = cam.rot90 ? cam.img_h : cam.img_w;
    const uint8_t* px = images + img_off + ((int64_t)r * stored_w + cc) * 3;
    black = (px[0] | px[1] | px[2]) == 0;
  }
  if (black && cam.black_mode == 1) ok = false;
  return ok;
}

// Binned path, step 1: ONE THREAD per point (the warp-per-point kernel spends a full warp instruction on every fp64 operation
// of at most 16 active lanes: at 2 M points its 1.8 ms are the fp64 pipe, not memory).  Winner camera, pixel, bin; the bin past
// the last camera's collects the points that get a zero row.  Histogram: one atomic per distinct bin of a warp, into one of
// PG_HIST_COPIES private copies (2 M atomics on the 48 lines of a single histogram would serialise in L2).
__global__ void __launch_bounds__(256)
    classify_points_kernel(const float* __restrict__ points, int64_t n, CamPack pack, int n_cam, const uint8_t* __restrict__ images,
                           int32_t* __restrict__ cam_of_point, int32_t* __restrict__ uv, int32_t* __restrict__ bin_of, int32_t* __restrict__ fuv,
                           int32_t* __restrict__ hist) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  const int n_bins = pack.cam[0].n_bins;
  int bin = n_bins - 1;
  if (live) {
    const double x = (double)points[i * 3 + 0], y = (double)points[i * 3 + 1], z = (double)points[i * 3 + 2];
    int found = -1, fu = 0, fv = 0;
    bool zero_feat = false;
    for (int c = 0; c < n_cam && found < 0; ++c) {   // first camera wins
      int u, v;
      bool black;
      if (project_one(pack.cam[c].c, images, pack.cam[c].img_off, x, y, z, u, v, black)) {
        found = c;
        fu = u;
        fv = v;
        zero_feat = black;
      }
    }
    if (cam_of_point) cam_of_point[i] = found;
    if (uv) {
      uv[2 * i] = found >= 0 ? fu : -1;
      uv[2 * i + 1] = found >= 0 ? fv : -1;
    }
    int packed = -1;
    if (found >= 0 && !zero_feat) {
      const vfmreg_camera& cam = pack.cam[found].c;
      int y0, y1, x0, x1;
      float wy, wx;
      cell_of(cam, fu, fv, y0, y1, x0, x1, wy, wx);
      bin = pack.cam[found].bin_base + y0 * cam.grid_w + x0;
      packed = (found << 28) | (fv << 14) | fu;   // 14 bits per pixel coordinate, 3 for the camera
    }
    bin_of[i] = bin;
    fuv[i] = packed;
  }
  const unsigned active = __ballot_sync(0xffffffffu, live);
  if (live) {
    const unsigned same = __match_any_sync(active, bin);
    if ((threadIdx.x & 31) == __ffs(same) - 1)
      atomicAdd(hist + (size_t)(blockIdx.x & (PG_HIST_COPIES - 1)) * n_bins + bin, __popc(same));
  }
}

__global__ void bin_offsets_kernel(int32_t* __restrict__ hist, int n_bins, int32_t* __restrict__ start) {
  // one CTA: sum of the private histograms, exclusive scan into `start`; copy 0 of the histogram becomes the per-bin cursor (zeroed)
  __shared__ int32_t carry;
  __shared__ int32_t wsum[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < n_bins; b0 += blockDim.x) {
    const int b = b0 + threadIdx.x;
    int32_t v = 0;
    if (b < n_bins)
      for (int c = 0; c < PG_HIST_COPIES; ++c) v += hist[(size_t)c * n_bins + b];
    int32_t incl = v;
    for (int off = 1; off < 32; off <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, incl, off);
      if ((threadIdx.x & 31) >= off) incl += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    int32_t base = carry;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) base += wsum[w];
    if (b < n_bins) {
      start[b] = base + incl - v;
      hist[b] = 0;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = base + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) start[n_bins] = carry;   // number of binned points
}

__global__ void __launch_bounds__(256)
    bin_scatter_kernel(const int32_t* __restrict__ bin_of, int64_t n, const int32_t* __restrict__ start, int32_t* __restrict__ cursor,
                       int32_t* __restrict__ order) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  const unsigned active = __ballot_sync(0xffffffffu, live);
  if (!live) return;
  const int bin = bin_of[i];
  // one atomic per distinct bin of the warp; the order inside a bin is irrelevant: rows are independent
  const unsigned same = __match_any_sync(active, bin);
  const int leader = __ffs(same) - 1, lane = threadIdx.x & 31;
  int pos = 0;
  if (lane == leader) pos = atomicAdd(cursor + bin, __popc(same));
  pos = __shfl_sync(same, pos, leader) + __popc(same & ((1u << lane) - 1));
  order[start[bin] + pos] = (int32_t)i;
}

// A warp takes 32 consecutive entries of the bin order.  Lane e prepares entry e (its four token-row offsets and blend
// weights); the warp then walks the entries, 128 channels (one float4 per lane) at a time, reloading the four rows only when
// the cell changes.
__global__ void __launch_bounds__(256)
    gather_binned_kernel(const int32_t* __restrict__ order, const int32_t* __restrict__ fuv, const int32_t* __restrict__ n_binned, CamPack pack,
                         const float* __restrict__ tokens, int d, float* __restrict__ desc) {
  const int lane = threadIdx.x & 31;
  const int64_t n = *n_binned;   // written by bin_offsets_kernel; the grid is sized for every point
  const int64_t base = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
  if (base >= n) return;
  const int count = (int)min((int64_t)32, n - base);
  int my_i = -1, o00 = -1, o01 = 0, o10 = 0, o11 = 0;
  float my_wy = 0.f, my_wx = 0.f;
  if (lane < count) {
    my_i = order[base + lane];
    const int f = fuv[my_i];
    if (f >= 0) {
      const int c = f >> 28, fv = (f >> 14) & 0x3fff, fu = f & 0x3fff;
      const vfmreg_camera& cam = pack.cam[c].c;
      int y0, y1, x0, x1;
      cell_of(cam, fu, fv, y0, y1, x0, x1, my_wy, my_wx);
      const int g = (int)pack.cam[c].tok_off;
      o00 = g + (y0 * cam.grid_w + x0) * d;
      o01 = g + (y0 * cam.grid_w + x1) * d;
      o10 = g + (y1 * cam.grid_w + x0) * d;
      o11 = g + (y1 * cam.grid_w + x1) * d;
    }
  }
  for (int k = 4 * lane; k < d; k += 128) {
    int c00 = -2, c01 = -2, c10 = -2, c11 = -2;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, c = a, e = a;
    for (int j = 0; j < count; ++j) {
      const int i = __shfl_sync(0xffffffffu, my_i, j);
      const int p00 = __shfl_sync(0xffffffffu, o00, j);
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p00 >= 0) {   // warp-uniform
        const int p01 = __shfl_sync(0xffffffffu, o01, j), p10 = __shfl_sync(0xffffffffu, o10, j), p11 = __shfl_sync(0xffffffffu, o11, j);
        const float wy = __shfl_sync(0xffffffffu, my_wy, j), wx = __shfl_sync(0xffffffffu, my_wx, j);
        if (p00 != c00 || p01 != c01 || p10 != c10 || p11 != c11) {
          a = __ldg(reinterpret_cast<const float4*>(tokens + p00 + k));
          b = __ldg(reinterpret_cast<const float4*>(tokens + p01 + k));
          c = __ldg(reinterpret_cast<const float4*>(tokens + p10 + k));
          e = __ldg(reinterpret_cast<const float4*>(tokens + p11 + k));
          c00 = p00; c01 = p01; c10 = p10; c11 = p11;
        }
        const float w0y = __fsub_rn(1.f, wy), w0x = __fsub_rn(1.f, wx);
#define VFM_BLEND(a_, b_, c_, d_) \
  __fadd_rn(__fmul_rn(w0y, __fadd_rn(__fmul_rn(w0x, a_), __fmul_rn(wx, b_))), __fmul_rn(wy, __fadd_rn(__fmul_rn(w0x, c_), __fmul_rn(wx, d_))))
        o.x = VFM_BLEND(a.x, b.x, c.x, e.x);
        o.y = VFM_BLEND(a.y, b.y, c.y, e.y);
        o.z = VFM_BLEND(a.z, b.z, c.z, e.z);
        o.w = VFM_BLEND(a.w, b.w, c.w, e.w);
#undef VFM_BLEND
      }
      __stcs(reinterpret_cast<float4*>(desc + (int64_t)i * d + k), o);
    }
  }
}

}  // namespace vfm

using namespace vfm;

// VFMREG_PROJECT_BINNED=0 keeps large clouds on the one-kernel path (A/B measurements, tools/bench_kernels.py project)
static const bool g_no_binned = [] { const char* e = getenv("VFMREG_PROJECT_BINNED"); return e && e[0] == '0'; }();

extern "C" int vfmreg_project_gather(vfmreg_ctx* ctx, const float* points, int64_t n, const vfmreg_camera* cams,
                                     int32_t n_cam, const float* tokens, const int64_t* token_offsets,
                                     const uint8_t* images, const int64_t* image_offsets, int32_t d, float* desc,
                                     int32_t* cam_of_point, int32_t* uv) {
  VFM_CHECK_ARG(ctx && points && cams && tokens && token_offsets && desc, "project_gather: null pointer");
  VFM_CHECK_ARG(n >= 0 && d > 0, "project_gather: bad sizes");
  VFM_CHECK_ARG(n_cam > 0 && n_cam <= MAX_CAMS, "project_gather: between 1 and %d cameras supported, got %d", MAX_CAMS, n_cam);
  VFM_CHECK_ARG(!images || image_offsets, "project_gather: images without offsets");
  VFM_CHECK_ARG((reinterpret_cast<uintptr_t>(tokens) & 15) == 0 && (reinterpret_cast<uintptr_t>(desc) & 15) == 0,
                "project_gather: tokens/desc must be 16-byte aligned");
  CamPack pack;
  memset(&pack, 0, sizeof(pack));
  for (int c = 0; c < n_cam; ++c) {
    pack.cam[c].c = cams[c];
    VFM_CHECK_ARG(cams[c].grid_h > 0 && cams[c].grid_w > 0 && cams[c].img_h > 0 && cams[c].img_w > 0 && cams[c].subsample > 0,
                  "project_gather: camera %d has an empty grid/image", c);
    VFM_CHECK_ARG((token_offsets[c] & 3) == 0 || (d & 3) != 0, "project_gather: token offset of camera %d not 16-byte aligned", c);
    pack.cam[c].tok_off = token_offsets[c];
    pack.cam[c].img_off = images ? image_offsets[c] : 0;
  }
  if (n == 0) return VFMREG_OK;
  VFM_CUDA(cudaSetDevice(ctx->device));
  group_begin(ctx, GROUP_PROJECT);
  // binned path: large clouds, 128-bit rows, pixel coordinates that fit 14 bits and token offsets that fit an int
  bool binned = n >= PG_BINNED_MIN && n < (1LL << 31) && (d & 3) == 0 && !g_no_binned;
  int n_bins = 0;
  int64_t tok_end = 0;
  for (int c = 0; c < n_cam; ++c) {
    pack.cam[c].bin_base = n_bins;
    n_bins += cams[c].grid_h * cams[c].grid_w;
    binned = binned && cams[c].crop_w < (1 << 14) && cams[c].crop_h < (1 << 14);
    const int64_t end = token_offsets[c] + (int64_t)cams[c].grid_h * cams[c].grid_w * d;
    tok_end = end > tok_end ? end : tok_end;
  }
  n_bins += 1;   // the last bin: points that get a zero row (unseen, or seen on a black pixel that claims them)
  for (int c = 0; c < MAX_CAMS; ++c) pack.cam[c].n_bins = n_bins;
  binned = binned && tok_end < (1LL << 31) && n_cam <= 8;   // the camera index has 3 bits (+ sign) in the packed word
  if (!binned) {
    project_gather_kernel<<<ceil_div(n, 8), 256, 0, ctx->stream>>>(points, n, pack, n_cam, tokens, images, d, desc, cam_of_point, uv);
    VFM_TRY(launch_check(ctx, "project_gather_kernel"));
    group_end(ctx, GROUP_PROJECT, 1);
    return VFMREG_OK;
  }
  arena_reset(ctx);
  VFM_TRY(arena_reserve(ctx, 3 * arena_bytes((size_t)n, 4) + arena_bytes((size_t)n_bins * PG_HIST_COPIES, 4) + arena_bytes((size_t)n_bins + 1, 4) + 4096));
  int32_t* bin_of = arena_take<int32_t>(ctx, (size_t)n);
  int32_t* fuv = arena_take<int32_t>(ctx, (size_t)n);
  int32_t* order = arena_take<int32_t>(ctx, (size_t)n);
  int32_t* hist = arena_take<int32_t>(ctx, (size_t)n_bins * PG_HIST_COPIES);
  int32_t* start = arena_take<int32_t>(ctx, (size_t)n_bins + 1);
  VFM_CHECK_ARG(bin_of && fuv && order && hist && start, "project_gather: scratch arena too small");
  VFM_CUDA(cudaMemsetAsync(hist, 0, (size_t)n_bins * PG_HIST_COPIES * 4, ctx->stream));
  classify_points_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(points, n, pack, n_cam, images, cam_of_point, uv, bin_of, fuv, hist);
  VFM_TRY(launch_check(ctx, "classify_points_kernel"));
  bin_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(hist, n_bins, start);
  VFM_TRY(launch_check(ctx, "bin_offsets_kernel"));
  bin_scatter_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(bin_of, n, start, hist, order);
  VFM_TRY(launch_check(ctx, "bin_scatter_kernel"));
  gather_binned_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(order, fuv, start + n_bins, pack, tokens, d, desc);
  VFM_TRY(launch_check(ctx, "gather_binned_kernel"));
  group_end(ctx, GROUP_PROJECT, 4);
  return VFMREG_OK;
}
