// DINOv2 ViT forward: model object, weight upload and the per-forward kernel sequence (C ABI: vfmreg_vit_*).
// Replaces ImageFeatureGenerator.get_image_features' GPU work (reference image_features.py:95-110: transform +
// self.model.model(x)), producing the channel-normalised patch-token grid (B, 16, patch_w, C) in fp32.
// GEMM operands are bf16 (weights converted once at upload), accumulation and the residual stream are fp32.
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"
#include "vit.cuh"

using namespace vfm;

namespace {

struct Layer {
  float *ln1_g, *ln1_b, *qkv_b, *proj_b, *ls1, *ln2_g, *ln2_b, *fc1_b, *fc2_b, *ls2;
  __nv_bfloat16 *qkv_w, *proj_w, *fc1_w, *fc2_w;
  CUtensorMap m_qkv, m_proj, m_fc1, m_fc2;   // 128-row boxes (vit_weight_map)
};

}  // namespace

struct vfmreg_vit {
  vfmreg_ctx* ctx;
  vfmreg_vit_config cfg;
  int kp;  // padded patch-embedding K (3 * patch^2 -> multiple of 64)
  __nv_bfloat16* pe_w = nullptr;
  CUtensorMap m_pe;
  float *pe_b = nullptr, *cls = nullptr, *norm_g = nullptr, *norm_b = nullptr, *cn_g = nullptr, *cn_b = nullptr;
  std::vector<Layer> layers;
  std::map<long long, float*> pos;  // (gh << 20 | gw) -> device (1 + gh*gw, W)
  std::vector<void*> allocs;
  // activations (grown on demand)
  int cap_rows = 0;
  float* x = nullptr;
  __nv_bfloat16 *xn = nullptr, *qkv = nullptr, *ao = nullptr, *hbuf = nullptr, *patches = nullptr;
  // per GEMM of a layer: how it is cut for the current row count (vit_gemm_plan) and the token maps with matching boxes
  GemmPlan p_pe{}, p_qkv{}, p_proj{}, p_fc1{}, p_fc2{};
  TokenMaps m_xn_qkv, m_xn_fc1, m_ao, m_h, m_patches;
  CUtensorMap m_qkv_attn;   // the QKV matrix for the attention kernel: boxes of 64 rows x 64 columns
  float* ws = nullptr;   // fp32 partial sums of proj / fc2: [split][rows][width]
  static constexpr int MAX_SPLIT = 4;
  int map_rows = -1, map_prow = -1;
  std::map<std::string, bool> loaded;
  // CUDA-graph replay of the per-forward kernel sequence (launch-bound at small batch): one graph per (b, h, w), captured on
  // the second call with that shape; images / tokens go through fixed staging buffers so the graph has no changing pointers
  struct Graph {
    cudaGraphExec_t exec = nullptr;
    int calls = 0;
    int launches = 0;
  };
  std::map<long long, Graph> graphs;
  uint8_t* img_stage = nullptr;
  float* tok_stage = nullptr;
  size_t img_stage_cap = 0, tok_stage_cap = 0;
  int use_graphs = 1;
  bool fused_resid = true;   // VFMREG_VIT_RESID=0 at creation: always partial sums + fold (residual_gemm)
  cudaStream_t cap_stream = nullptr;   // capture happens on a private stream (the legacy default stream cannot capture)
};

namespace {

template <typename T>
int dev_alloc(vfmreg_vit* v, T** p, size_t count) {
  cudaError_t e = cudaMalloc(p, count * sizeof(T));
  if (e != cudaSuccess) {
    set_error("vit: cudaMalloc(%zu) failed: %s", count * sizeof(T), cudaGetErrorString(e));
    return VFMREG_ERR_ALLOC;
  }
  v->allocs.push_back(*p);
  return VFMREG_OK;
}

__global__ void f32_to_bf16_pad_kernel(const float* __restrict__ src, int rows, int cols, int cols_pad, __nv_bfloat16* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * cols_pad) return;
  const int r = (int)(i / cols_pad), c = (int)(i % cols_pad);
  dst[i] = __float2bfloat16(c < cols ? src[(long long)r * cols + c] : 0.f);
}

int upload_f32(vfmreg_vit* v, float* dst, const float* host, size_t count) {
  VFM_CUDA(cudaMemcpy(dst, host, count * sizeof(float), cudaMemcpyHostToDevice));
  return VFMREG_OK;
}

int upload_bf16(vfmreg_vit* v, __nv_bfloat16* dst, const float* host, int rows, int cols, int cols_pad) {
  float* tmp = nullptr;
  VFM_CUDA(cudaMalloc(&tmp, (size_t)rows * cols * sizeof(float)));
  cudaError_t e = cudaMemcpy(tmp, host, (size_t)rows * cols * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    const long long total = (long long)rows * cols_pad;
    f32_to_bf16_pad_kernel<<<ceil_div(total, 256), 256, 0, v->ctx->stream>>>(tmp, rows, cols, cols_pad, dst);
    e = cudaStreamSynchronize(v->ctx->stream);
  }
  cudaFree(tmp);
  if (e != cudaSuccess) {
    set_error("vit: weight upload failed: %s", cudaGetErrorString(e));
    return VFMREG_ERR_CUDA;
  }
  v->ctx->launches += 1;
  return VFMREG_OK;
}

int ensure_activations(vfmreg_vit* v, int rows, int prows) {
  const int w = v->cfg.width, md = v->cfg.mlp_dim;
  if (rows > v->cap_rows) {
    VFM_CUDA(cudaStreamSynchronize(v->ctx->stream));
    for (void* p : {(void*)v->x, (void*)v->xn, (void*)v->qkv, (void*)v->ao, (void*)v->hbuf, (void*)v->patches, (void*)v->ws})
      if (p) cudaFree(p);
    for (auto& kv : v->graphs)   // captured graphs hold the old activation pointers
      if (kv.second.exec) {
        cudaGraphExecDestroy(kv.second.exec);
        kv.second.exec = nullptr;
      }
    const int cap = rows + rows / 4;
    VFM_CUDA(cudaMalloc(&v->x, (size_t)cap * w * sizeof(float)));
    VFM_CUDA(cudaMalloc(&v->xn, (size_t)cap * w * 2));
    VFM_CUDA(cudaMalloc(&v->qkv, (size_t)cap * 3 * w * 2));
    VFM_CUDA(cudaMalloc(&v->ao, (size_t)cap * w * 2));
    VFM_CUDA(cudaMalloc(&v->hbuf, (size_t)cap * md * 2));
    VFM_CUDA(cudaMalloc(&v->patches, (size_t)cap * v->kp * 2));
    VFM_CUDA(cudaMalloc(&v->ws, (size_t)vfmreg_vit::MAX_SPLIT * cap * w * sizeof(float)));
    v->cap_rows = cap;
    v->map_rows = -1;
  }
  if (rows != v->map_rows || prows != v->map_prow) {
    v->p_pe = vit_gemm_plan(v->ctx, prows, w, v->kp, 1, "pe");
    v->p_qkv = vit_gemm_plan(v->ctx, rows, 3 * w, w, 1, "qkv");
    v->p_proj = vit_gemm_plan(v->ctx, rows, w, w, vfmreg_vit::MAX_SPLIT, "proj");
    v->p_fc1 = vit_gemm_plan(v->ctx, rows, md, w, 1, "fc1");
    v->p_fc2 = vit_gemm_plan(v->ctx, rows, w, md, vfmreg_vit::MAX_SPLIT, "fc2");
    if (getenv("VFMREG_VIT_VERBOSE"))   // tuning aid: how the cost model cut the five GEMMs for this row count
      for (const auto& pr : {std::make_pair("pe", v->p_pe), std::make_pair("qkv", v->p_qkv), std::make_pair("proj", v->p_proj),
                             std::make_pair("fc1", v->p_fc1), std::make_pair("fc2", v->p_fc2)})
        fprintf(stderr, "vit plan rows=%d %s: nt=%d split=%d tail=%d units=%d\n", rows, pr.first, pr.second.nt, pr.second.split,
                pr.second.tail_w, pr.second.units);
    VFM_TRY(vit_token_maps(&v->m_xn_qkv, v->xn, rows, w, v->p_qkv));
    VFM_TRY(vit_token_maps(&v->m_xn_fc1, v->xn, rows, w, v->p_fc1));
    VFM_TRY(vit_token_maps(&v->m_ao, v->ao, rows, w, v->p_proj));
    VFM_TRY(vit_token_maps(&v->m_h, v->hbuf, rows, md, v->p_fc2));
    VFM_TRY(vit_token_maps(&v->m_patches, v->patches, prows, v->kp, v->p_pe));
    VFM_TRY(make_tmap_16bit(&v->m_qkv_attn, v->qkv, rows, 3 * w, 3 * w, 64, true));
    v->map_rows = rows;
    v->map_prow = prows;
  }
  return VFMREG_OK;
}

}  // namespace

extern "C" {

int vfmreg_vit_create(vfmreg_ctx* ctx, const vfmreg_vit_config* cfg, vfmreg_vit** out) {
  VFM_CHECK_ARG(ctx && cfg && out, "vit_create: null pointer");
  VFM_CHECK_ARG(cfg->depth > 0 && cfg->heads > 0 && cfg->width == cfg->heads * 64, "vit_create: width must be heads * 64");
  VFM_CHECK_ARG(cfg->width % 128 == 0 && cfg->width <= 1024 && cfg->mlp_dim % 64 == 0, "vit_create: unsupported width/mlp_dim");
  VFM_CHECK_ARG(cfg->mlp_dim % 128 == 0, "vit_create: mlp_dim must be a multiple of 128");
  VFM_CHECK_ARG(cfg->patch > 0 && cfg->patch_h > 0, "vit_create: bad patch geometry");
  VFM_CUDA(cudaSetDevice(ctx->device));
  vfmreg_vit* v = new vfmreg_vit();
  if (const char* e = getenv("VFMREG_VIT_RESID")) v->fused_resid = e[0] != '0';
  v->ctx = ctx;
  v->cfg = *cfg;
  const int w = cfg->width, md = cfg->mlp_dim;
  v->kp = (3 * cfg->patch * cfg->patch + 63) / 64 * 64;
  int rc = VFMREG_OK;
  auto A = [&](auto** p, size_t n) { if (rc == VFMREG_OK) rc = dev_alloc(v, p, n); };
  A(&v->pe_w, (size_t)w * v->kp);
  A(&v->pe_b, w); A(&v->cls, w); A(&v->norm_g, w); A(&v->norm_b, w); A(&v->cn_g, w); A(&v->cn_b, w);
  v->layers.resize(cfg->depth);
  for (Layer& l : v->layers) {
    A(&l.ln1_g, w); A(&l.ln1_b, w); A(&l.qkv_b, 3 * w); A(&l.proj_b, w); A(&l.ls1, w);
    A(&l.ln2_g, w); A(&l.ln2_b, w); A(&l.fc1_b, md); A(&l.fc2_b, w); A(&l.ls2, w);
    A(&l.qkv_w, (size_t)3 * w * w); A(&l.proj_w, (size_t)w * w); A(&l.fc1_w, (size_t)md * w); A(&l.fc2_w, (size_t)w * md);
  }
  if (rc == VFMREG_OK) rc = vit_weight_map(&v->m_pe, v->pe_w, w, v->kp);
  for (Layer& l : v->layers) {
    if (rc != VFMREG_OK) break;
    rc = vit_weight_map(&l.m_qkv, l.qkv_w, 3 * w, w);
    if (rc == VFMREG_OK) rc = vit_weight_map(&l.m_proj, l.proj_w, w, w);
    if (rc == VFMREG_OK) rc = vit_weight_map(&l.m_fc1, l.fc1_w, md, w);
    if (rc == VFMREG_OK) rc = vit_weight_map(&l.m_fc2, l.fc2_w, w, md);
  }
  if (rc != VFMREG_OK) {
    vfmreg_vit_destroy(v);
    return rc;
  }
  // identity ChannelNorm affine by default
  std::vector<float> ones(w, 1.f), zeros(w, 0.f);
  cudaMemcpy(v->cn_g, ones.data(), w * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(v->cn_b, zeros.data(), w * 4, cudaMemcpyHostToDevice);
  *out = v;
  return VFMREG_OK;
}

void vfmreg_vit_destroy(vfmreg_vit* v) {
  if (!v) return;
  cudaSetDevice(v->ctx->device);
  cudaStreamSynchronize(v->ctx->stream);
  for (void* p : v->allocs) cudaFree(p);
  for (auto& kv : v->pos) cudaFree(kv.second);
  for (auto& kv : v->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  if (v->img_stage) cudaFree(v->img_stage);
  if (v->tok_stage) cudaFree(v->tok_stage);
  if (v->cap_stream) cudaStreamDestroy(v->cap_stream);
  for (void* p : {(void*)v->x, (void*)v->xn, (void*)v->qkv, (void*)v->ao, (void*)v->hbuf, (void*)v->patches, (void*)v->ws})
    if (p) cudaFree(p);
  delete v;
}

// Names follow the dinov2 hub state dict: patch_embed.proj.weight|bias, cls_token, norm.weight|bias,
// channel_norm.weight|bias, blocks.<l>.{norm1,norm2}.{weight,bias}, blocks.<l>.attn.{qkv,proj}.{weight,bias},
// blocks.<l>.{ls1,ls2}.gamma, blocks.<l>.mlp.{fc1,fc2}.{weight,bias}.  host: float32, row-major, `count` elements.
int vfmreg_vit_set_weight(vfmreg_vit* v, const char* name, const float* host, int64_t count) {
  VFM_CHECK_ARG(v && name && host, "vit_set_weight: null pointer");
  VFM_CUDA(cudaSetDevice(v->ctx->device));
  const int w = v->cfg.width, md = v->cfg.mlp_dim, pk = 3 * v->cfg.patch * v->cfg.patch;
  const std::string n(name);
  auto need = [&](int64_t want) {
    if (count != want) {
      set_error("vit_set_weight(%s): expected %lld elements, got %lld", name, (long long)want, (long long)count);
      return false;
    }
    return true;
  };
  int rc = VFMREG_ERR_ARG;
  if (n == "patch_embed.proj.weight") { if (need((int64_t)w * pk)) rc = upload_bf16(v, v->pe_w, host, w, pk, v->kp); }
  else if (n == "patch_embed.proj.bias") { if (need(w)) rc = upload_f32(v, v->pe_b, host, w); }
  else if (n == "cls_token") { if (need(w)) rc = upload_f32(v, v->cls, host, w); }
  else if (n == "norm.weight") { if (need(w)) rc = upload_f32(v, v->norm_g, host, w); }
  else if (n == "norm.bias") { if (need(w)) rc = upload_f32(v, v->norm_b, host, w); }
  else if (n == "channel_norm.weight") { if (need(w)) rc = upload_f32(v, v->cn_g, host, w); }
  else if (n == "channel_norm.bias") { if (need(w)) rc = upload_f32(v, v->cn_b, host, w); }
  else if (n.rfind("blocks.", 0) == 0) {
    const size_t dot = n.find('.', 7);
    const int li = atoi(n.substr(7, dot - 7).c_str());
    if (dot == std::string::npos || li < 0 || li >= v->cfg.depth) {
      set_error("vit_set_weight: bad layer in %s", name);
      return VFMREG_ERR_ARG;
    }
    Layer& l = v->layers[li];
    const std::string k = n.substr(dot + 1);
    if (k == "norm1.weight") { if (need(w)) rc = upload_f32(v, l.ln1_g, host, w); }
    else if (k == "norm1.bias") { if (need(w)) rc = upload_f32(v, l.ln1_b, host, w); }
    else if (k == "norm2.weight") { if (need(w)) rc = upload_f32(v, l.ln2_g, host, w); }
    else if (k == "norm2.bias") { if (need(w)) rc = upload_f32(v, l.ln2_b, host, w); }
    else if (k == "attn.qkv.weight") { if (need((int64_t)3 * w * w)) rc = upload_bf16(v, l.qkv_w, host, 3 * w, w, w); }
    else if (k == "attn.qkv.bias") { if (need(3 * w)) rc = upload_f32(v, l.qkv_b, host, 3 * w); }
    else if (k == "attn.proj.weight") { if (need((int64_t)w * w)) rc = upload_bf16(v, l.proj_w, host, w, w, w); }
    else if (k == "attn.proj.bias") { if (need(w)) rc = upload_f32(v, l.proj_b, host, w); }
    else if (k == "ls1.gamma") { if (need(w)) rc = upload_f32(v, l.ls1, host, w); }
    else if (k == "ls2.gamma") { if (need(w)) rc = upload_f32(v, l.ls2, host, w); }
    else if (k == "mlp.fc1.weight") { if (need((int64_t)md * w)) rc = upload_bf16(v, l.fc1_w, host, md, w, w); }
    else if (k == "mlp.fc1.bias") { if (need(md)) rc = upload_f32(v, l.fc1_b, host, md); }
    else if (k == "mlp.fc2.weight") { if (need((int64_t)w * md)) rc = upload_bf16(v, l.fc2_w, host, w, md, md); }
    else if (k == "mlp.fc2.bias") { if (need(w)) rc = upload_f32(v, l.fc2_b, host, w); }
    else set_error("vit_set_weight: unknown tensor %s", name);
  } else {
    set_error("vit_set_weight: unknown tensor %s", name);
  }
  if (rc == VFMREG_OK) v->loaded[n] = true;
  return rc;
}

// Position embedding already interpolated to the (grid_h, grid_w) patch grid: (1 + grid_h*grid_w, width) float32, CLS first.
int vfmreg_vit_set_pos_embed(vfmreg_vit* v, int32_t grid_h, int32_t grid_w, const float* host) {
  VFM_CHECK_ARG(v && host && grid_h > 0 && grid_w > 0, "vit_set_pos_embed: bad arguments");
  VFM_CUDA(cudaSetDevice(v->ctx->device));
  const long long key = ((long long)grid_h << 20) | grid_w;
  const size_t count = (size_t)(1 + grid_h * grid_w) * v->cfg.width;
  float* d = nullptr;
  auto it = v->pos.find(key);
  if (it != v->pos.end()) {
    d = it->second;
  } else {
    VFM_CUDA(cudaMalloc(&d, count * sizeof(float)));
    v->pos[key] = d;
  }
  VFM_CUDA(cudaMemcpy(d, host, count * sizeof(float), cudaMemcpyHostToDevice));
  return VFMREG_OK;
}

int vfmreg_vit_grid(const vfmreg_vit* v, int32_t img_h, int32_t img_w, int32_t* grid_h, int32_t* grid_w) {
  VFM_CHECK_ARG(v && grid_h && grid_w && img_h > 0 && img_w > 0, "vit_grid: bad arguments");
  // create_transform_ (image_features.py:67-69): scale = (patch * patch_h) / H; patch_w = int(scale * W / patch), in doubles
  const double scale = (double)(v->cfg.patch * v->cfg.patch_h) / (double)img_h;
  *grid_h = v->cfg.patch_h;
  *grid_w = (int32_t)(scale * (double)img_w / (double)v->cfg.patch);
  return VFMREG_OK;
}

// One residual branch, x += scale * (X W^T + bias).  When the plan keeps K whole the add rides in the GEMM's epilogue and
// nothing is pending; with K split over several units (few images: it takes the splits to fill the GPU) the units write fp32
// partial sums and the next normalisation kernel folds them into x in a fixed order.  VFMREG_VIT_RESID=0: always the latter.
static int residual_gemm(vfmreg_vit* v, int rows, int n, int k, const CUtensorMap& m_w, const TokenMaps& m_x, const GemmPlan& plan,
                         const float* bias, const float* scale, const void* pf_ptr, size_t pf_bytes, Residual* pending) {
  GemmEpilogue ep{};
  ep.m = rows; ep.n = n; ep.k = k; ep.ldo = n;
  if (pf_ptr) { ep.pf_ptr = (const char*)pf_ptr; ep.pf_bytes = pf_bytes; }
  if (v->fused_resid && plan.split == 1) {
    ep.x = v->x; ep.bias = bias; ep.scale = scale;
    *pending = Residual{};
    return vit_gemm(v->ctx, EPI_F32_RESID, m_w, m_x, plan, ep);
  }
  ep.x = v->ws;
  *pending = Residual{v->ws, plan.split, bias, scale};
  return vit_gemm(v->ctx, EPI_F32_PARTIAL, m_w, m_x, plan, ep);
}

static int vit_enqueue(vfmreg_vit* v, const uint8_t* images, int b, int img_h, int img_w, int gh, int gw, const float* pos,
                       float* tokens) {
  vfmreg_ctx* ctx = v->ctx;
  const int w = v->cfg.width, md = v->cfg.mlp_dim, np = gh * gw, t = np + 1;
  const int rows = b * t, prows = b * np;
  const float ms[6] = {v->cfg.mean[0], v->cfg.mean[1], v->cfg.mean[2], v->cfg.std[0], v->cfg.std[1], v->cfg.std[2]};
  VFM_TRY(vit_preprocess(ctx, images, b, img_h, img_w, gh, gw, v->cfg.patch, ms, v->patches, v->kp, v->x, v->cls, pos, w));
  GemmEpilogue ep{};
  ep.m = prows; ep.n = w; ep.k = v->kp; ep.np = np; ep.ldo = w; ep.bias = v->pe_b; ep.pos = pos; ep.x = v->x;
  VFM_TRY(vit_gemm(ctx, EPI_F32_PATCH, v->m_pe, v->m_patches, v->p_pe, ep));
  Residual pending{};   // the residual branch of the last proj / fc2 GEMM, folded into the next normalisation kernel
  for (size_t li = 0; li < v->layers.size(); ++li) {
    Layer& l = v->layers[li];
    // every GEMM / attention kernel pulls the weights of a GEMM further down the chain into L2 while it runs
    // (small batches only: with thousands of tokens a GEMM runs for tens of microseconds and its weights are a small share
    // of its traffic)
    const bool pf = rows <= 4096;
    const Layer* nxt = pf && li + 1 < v->layers.size() ? &v->layers[li + 1] : nullptr;
    VFM_TRY(vit_layernorm_bf16(ctx, v->x, rows, w, pending, l.ln1_g, l.ln1_b, v->cfg.ln_eps, v->xn));
    ep = GemmEpilogue{};
    ep.m = rows; ep.n = 3 * w; ep.k = w; ep.ldo = 3LL * w; ep.bias = l.qkv_b; ep.out_bf16 = v->qkv;
    if (pf) { ep.pf_ptr = (const char*)l.proj_w; ep.pf_bytes = (size_t)w * w * 2; }
    VFM_TRY(vit_gemm(ctx, EPI_BF16_BIAS, l.m_qkv, v->m_xn_qkv, v->p_qkv, ep));
    if (vit_attention_tc_supported(t))
      VFM_TRY(vit_attention_tc(ctx, v->m_qkv_attn, b, t, v->cfg.heads, w, v->ao, pf ? l.fc1_w : nullptr, (size_t)md * w * 2));
    else
      VFM_TRY(vit_attention(ctx, v->qkv, b, t, v->cfg.heads, w, v->ao));
    VFM_TRY(residual_gemm(v, rows, w, w, l.m_proj, v->m_ao, v->p_proj, l.proj_b, l.ls1, nullptr, 0, &pending));
    VFM_TRY(vit_layernorm_bf16(ctx, v->x, rows, w, pending, l.ln2_g, l.ln2_b, v->cfg.ln_eps, v->xn));
    ep = GemmEpilogue{};
    ep.m = rows; ep.n = md; ep.k = w; ep.ldo = md; ep.bias = l.fc1_b; ep.out_bf16 = v->hbuf;
    if (pf) { ep.pf_ptr = (const char*)l.fc2_w; ep.pf_bytes = (size_t)w * md * 2; }
    VFM_TRY(vit_gemm(ctx, EPI_BF16_BIAS_GELU, l.m_fc1, v->m_xn_fc1, v->p_fc1, ep));
    VFM_TRY(residual_gemm(v, rows, w, md, l.m_fc2, v->m_h, v->p_fc2, l.fc2_b, l.ls2, nxt ? nxt->qkv_w : nullptr, (size_t)3 * w * w * 2,
                          &pending));
  }
  return vit_final_norm(ctx, v->x, b, t, w, pending, v->norm_g, v->norm_b, v->cfg.ln_eps, v->cn_g, v->cn_b, v->cfg.cn_eps, v->cfg.channel_norm,
                        tokens);
}

int vfmreg_vit_set_graphs(vfmreg_vit* v, int32_t on) {
  VFM_CHECK_ARG(v, "vit_set_graphs: null pointer");
  v->use_graphs = on;
  return VFMREG_OK;
}

int vfmreg_vit_forward(vfmreg_vit* v, const uint8_t* images, int32_t b, int32_t img_h, int32_t img_w, float* tokens) {
  VFM_CHECK_ARG(v && images && tokens && b > 0, "vit_forward: bad arguments");
  vfmreg_ctx* ctx = v->ctx;
  VFM_CUDA(cudaSetDevice(ctx->device));
  int32_t gh, gw;
  VFM_TRY(vfmreg_vit_grid(v, img_h, img_w, &gh, &gw));
  VFM_CHECK_ARG(gw > 0, "vit_forward: image %dx%d is too narrow for one patch column", img_h, img_w);
  auto it = v->pos.find(((long long)gh << 20) | gw);
  VFM_CHECK_ARG(it != v->pos.end(), "vit_forward: no position embedding set for the %dx%d patch grid", gh, gw);
  const float* pos = it->second;
  const int np = gh * gw, t = np + 1;
  VFM_TRY(ensure_activations(v, b * t, b * np));
  const size_t img_bytes = (size_t)b * img_h * img_w * 3, tok_bytes = (size_t)b * np * v->cfg.width * sizeof(float);
  const long long key = ((long long)b << 40) | ((long long)img_h << 20) | img_w;
  vfmreg_vit::Graph& g = v->graphs[key];
  g.calls += 1;
  group_begin(ctx, GROUP_VIT);
  int rc = VFMREG_OK;
  if (!v->use_graphs || g.calls == 1) {
    // first call with this shape: plain launches (also sets kernel attributes, which cannot happen during capture)
    const int64_t l0 = ctx->launches;
    rc = vit_enqueue(v, images, b, img_h, img_w, gh, gw, pos, tokens);
    g.launches = (int)(ctx->launches - l0);
    group_end(ctx, GROUP_VIT, g.launches);
    return rc;
  }
  if (img_bytes > v->img_stage_cap || tok_bytes > v->tok_stage_cap) {
    // staging buffers are baked into captured graphs: growing them invalidates every graph
    VFM_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto& kv : v->graphs)
      if (kv.second.exec) {
        cudaGraphExecDestroy(kv.second.exec);
        kv.second.exec = nullptr;
      }
    if (v->img_stage) cudaFree(v->img_stage);
    if (v->tok_stage) cudaFree(v->tok_stage);
    v->img_stage_cap = img_bytes + img_bytes / 2;
    v->tok_stage_cap = tok_bytes + tok_bytes / 2;
    VFM_CUDA(cudaMalloc(&v->img_stage, v->img_stage_cap));
    VFM_CUDA(cudaMalloc(&v->tok_stage, v->tok_stage_cap));
  }
  if (!g.exec) {
    cudaGraph_t graph = nullptr;
    if (!v->cap_stream) VFM_CUDA(cudaStreamCreateWithFlags(&v->cap_stream, cudaStreamNonBlocking));
    const int timing = ctx->timing;
    cudaStream_t user_stream = ctx->stream;
    ctx->timing = 0;  // no event records inside the capture
    ctx->stream = v->cap_stream;
    cudaError_t e = cudaStreamBeginCapture(v->cap_stream, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
      rc = vit_enqueue(v, v->img_stage, b, img_h, img_w, gh, gw, pos, v->tok_stage);
      e = cudaStreamEndCapture(v->cap_stream, &graph);
    }
    ctx->stream = user_stream;
    ctx->timing = timing;
    if (rc != VFMREG_OK) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    VFM_CUDA(e);
    e = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    VFM_CUDA(e);
  }
  VFM_CUDA(cudaMemcpyAsync(v->img_stage, images, img_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  VFM_CUDA(cudaGraphLaunch(g.exec, ctx->stream));
  VFM_CUDA(cudaMemcpyAsync(tokens, v->tok_stage, tok_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  ctx->launches += g.launches;  // kernels replayed by the graph
  group_end(ctx, GROUP_VIT, g.launches);
  return VFMREG_OK;
}

}  // extern "C"
