// Point -> pixel projection + feature gather + first-camera-wins scatter in one pass.
// Replaces project_pcl_to_image (reference dataloader/nclt.py:311-366, dataloader/oxford_robotcar.py:330-363) and the
// gather / dedup / scatter of create_descriptors (prepare_scenes.py:57-104).  The full-resolution feature map the
// reference materialises (F.interpolate to (H, W, C), image_features.py:104-108: ~0.9 GB per NCLT image) is never built:
// the value that map would hold at the integer pixel is computed from the 4 surrounding tokens.
//
// One warp per point.  Lane c evaluates the projection (float64, as NumPy does) into camera c, a ballot picks the first
// camera that sees the point (first camera wins, prepare_scenes.py:97-101); the lanes then stride the channel
// dimension with 128-bit loads of the L2-resident token grid and 128-bit streaming stores of the descriptor row.
// HBM-bound: 12 B read + 4 d B written per point (unseen points are written as zeros, prepare_scenes.py:102-104).
#include "common.cuh"

namespace vfm {

constexpr int MAX_CAMS = 16;

struct CamDev {
  vfmreg_camera c;
  int64_t tok_off;
  int64_t img_off;
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

}  // namespace vfm

using namespace vfm;

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
  project_gather_kernel<<<ceil_div(n, 8), 256, 0, ctx->stream>>>(points, n, pack, n_cam, tokens, images, d, desc, cam_of_point, uv);
  VFM_TRY(launch_check(ctx, "project_gather_kernel"));
  group_end(ctx, GROUP_PROJECT, 1);
  return VFMREG_OK;
}
