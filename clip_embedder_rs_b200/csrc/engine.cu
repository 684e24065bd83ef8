#include "engine.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "attn_sm100.cuh"
#include "attn_short_sm100.cuh"
#include "gemm_sm100.cuh"
#include "conv_kernels.cuh"
#include "kernels.cuh"
#include "search.cuh"

namespace clipb200 {

#define RET_IF_ERR(expr)      \
  do {                        \
    Status s_ = (expr);       \
    if (!s_.ok()) return s_;  \
  } while (0)
#define CUDA_RET(expr, what)                      \
  do {                                            \
    Status s_ = Check((expr), what);              \
    if (!s_.ok()) return s_;                      \
  } while (0)

Status Engine::Check(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return Status::OK();
  return Status::Err(CLIPB200_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

static float half_to_float(uint16_t h) {
  const uint32_t sign = (h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1F, man = h & 0x3FF, bits;
  if (exp == 0) {
    if (man == 0) bits = sign;
    else {
      exp = 127 - 15 + 1;
      while (!(man & 0x400)) { man <<= 1; --exp; }
      bits = sign | (exp << 23) | ((man & 0x3FF) << 13);
    }
  } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
  else bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

// the corpus search (capi.cu / search.cu) reaches the GEMM through this plain function
cudaError_t gemm_bf16_f32out(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw, int M, int N,
                             int K, float* out, long long ldc, bool accumulate, cudaStream_t st) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  static std::atomic<unsigned long long> configured(0);  // bit per device
  if (dev < 64 && !(configured.load() & (1ull << dev))) {
    if (get_encode_tiled() == nullptr) return cudaErrorNotSupported;
    if ((e = gemm_configure_device()) != cudaSuccess) return e;
    configured.fetch_or(1ull << dev);
  }
  int sms = 0;
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
  GemmEpilogue ep;
  ep.out_f32 = out;
  ep.ldc = ldc;
  return gemm_bf16(A, lda, W, ldw, M, N, K, accumulate ? EPI_RESID : EPI_F32, ep, sms, st);
}

// ------------------------------------------------------------------------------------------------ create
Status Engine::Create(const std::string& onnx_path, int device, const clipb200_opts* opts, Engine** out) {
  Engine* e = new Engine();
  Status s = e->Init(onnx_path, device, opts);
  if (!s.ok()) {
    delete e;
    return s;
  }
  *out = e;
  return Status::OK();
}

Status Engine::DevAlloc(void** p, size_t bytes) {
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess)
    return Status::Err(CLIPB200_ERR_CUDA, "cudaMalloc(" + std::to_string(bytes) + "): " + cudaGetErrorString(e));
  dev_allocs_.push_back(*p);
  return Status::OK();
}

Status Engine::HostF32(const OnnxModel& m, const std::string& name, int64_t expect_numel, std::vector<float>* out) {
  const OnnxTensor* t = m.find(name);
  if (t == nullptr) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "initializer '" + name + "' not found in graph");
  if (expect_numel >= 0 && t->numel() != expect_numel)
    return Status::Err(CLIPB200_ERR_UNSUPPORTED, "initializer '" + name + "' has " + std::to_string(t->numel()) +
                                                     " elements, expected " + std::to_string(expect_numel));
  const int64_t n = t->numel();
  out->resize(static_cast<size_t>(n));
  if (t->data_type == 1) {
    memcpy(out->data(), t->data, static_cast<size_t>(n) * 4);
  } else if (t->data_type == 10) {
    const uint16_t* p = reinterpret_cast<const uint16_t*>(t->data);
    for (int64_t i = 0; i < n; ++i) (*out)[i] = half_to_float(p[i]);
  } else if (t->data_type == 16) {
    const uint16_t* p = reinterpret_cast<const uint16_t*>(t->data);
    for (int64_t i = 0; i < n; ++i) {
      const uint32_t bits = static_cast<uint32_t>(p[i]) << 16;
      memcpy(&(*out)[i], &bits, 4);
    }
  } else {
    return Status::Err(CLIPB200_ERR_UNSUPPORTED,
                       "initializer '" + name + "' has data_type " + std::to_string(t->data_type) + " (need f32/f16/bf16)");
  }
  if (t->transposed && t->dims.size() == 2) {   // canonical [r, c] alias whose bytes are [c, r] (onnx_graph.cc): hand out [r, c]
    const int64_t r = t->dims[0], c = t->dims[1];
    std::vector<float> tmp(out->size());
    for (int64_t i = 0; i < r; ++i)
      for (int64_t j = 0; j < c; ++j) tmp[static_cast<size_t>(i * c + j)] = (*out)[static_cast<size_t>(j * r + i)];
    out->swap(tmp);
  }
  return Status::OK();
}

Status Engine::UploadF32(const OnnxModel& m, const std::string& name, int64_t expect_numel, float** out) {
  const OnnxTensor* t = m.find(name);
  if (t == nullptr) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "initializer '" + name + "' not found in graph");
  const float* src = nullptr;
  std::vector<float> tmp;
  if (t->data_type == 1 && (expect_numel < 0 || t->numel() == expect_numel)) {
    src = reinterpret_cast<const float*>(t->data);
  } else {
    RET_IF_ERR(HostF32(m, name, expect_numel, &tmp));
    src = tmp.data();
  }
  const size_t bytes = static_cast<size_t>(t->numel()) * 4;
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(out), bytes));
  CUDA_RET(cudaMemcpy(*out, src, bytes, cudaMemcpyHostToDevice), "upload fp32 initializer");
  weight_bytes += static_cast<int64_t>(bytes);
  return Status::OK();
}

// Linear weight -> bf16 [N, ldk] (K contiguous, zero padded to a multiple of 8).  `transpose`: the initializer is
// stored [K, N] (open_clip `proj` / `text_projection` parameters that are applied as x @ P).
Status Engine::UploadLinear(const OnnxModel& m, const std::string& wname, const std::string& bname, int N, int K,
                            bool transpose, LinearW* out) {
  const OnnxTensor* t = m.find(wname);
  if (t == nullptr) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "initializer '" + wname + "' not found in graph");
  if (t->numel() != static_cast<int64_t>(N) * K)
    return Status::Err(CLIPB200_ERR_UNSUPPORTED, "initializer '" + wname + "' has " + std::to_string(t->numel()) +
                                                     " elements, expected " + std::to_string(N) + "x" + std::to_string(K));
  const float* src = nullptr;
  std::vector<float> tmp;
  if (t->data_type == 1) src = reinterpret_cast<const float*>(t->data);
  else {
    RET_IF_ERR(HostF32(m, wname, -1, &tmp));
    src = tmp.data();
  }
  if (t->transposed) transpose = !transpose;  // canonical alias of a pre-transposed MatMul operand (onnx_graph.cc)
  const int ldk = (K + 7) & ~7;
  float* staging = nullptr;
  const size_t fbytes = static_cast<size_t>(N) * K * 4;
  CUDA_RET(cudaMalloc(&staging, fbytes), "staging alloc");
  cudaError_t ce = cudaMemcpy(staging, src, fbytes, cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) {
    cudaFree(staging);
    return Check(ce, "upload weight");
  }
  Status s = DevAlloc(reinterpret_cast<void**>(&out->w), static_cast<size_t>(N) * ldk * 2);
  if (s.ok()) s = Check(launch_convert_f32_bf16(staging, N, K, ldk, transpose, out->w, 0), "convert weight");
  if (s.ok()) s = Check(cudaDeviceSynchronize(), "convert weight sync");
  cudaFree(staging);
  RET_IF_ERR(s);
  weight_bytes += static_cast<int64_t>(N) * ldk * 2;
  out->N = N;
  out->K = K;
  out->ldk = ldk;
  out->b = nullptr;
  if (!bname.empty()) RET_IF_ERR(UploadF32(m, bname, N, &out->b));
  return Status::OK();
}

Status Engine::UploadHostF32(const float* src, size_t n, float** out) {
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(out), n * 4));
  CUDA_RET(cudaMemcpy(*out, src, n * 4, cudaMemcpyHostToDevice), "upload fp32 tensor");
  weight_bytes += static_cast<int64_t>(n) * 4;
  return Status::OK();
}

Status Engine::UploadLinearFromHost(const float* w, int N, int K, const float* bias_or_null, LinearW* out) {
  const int ldk = (K + 7) & ~7;
  float* staging = nullptr;
  const size_t fbytes = static_cast<size_t>(N) * K * 4;
  CUDA_RET(cudaMalloc(&staging, fbytes), "staging alloc");
  cudaError_t ce = cudaMemcpy(staging, w, fbytes, cudaMemcpyHostToDevice);
  Status s = Check(ce, "upload weight");
  if (s.ok()) s = DevAlloc(reinterpret_cast<void**>(&out->w), static_cast<size_t>(N) * ldk * 2);
  if (s.ok()) s = Check(launch_convert_f32_bf16(staging, N, K, ldk, false, out->w, 0), "convert weight");
  if (s.ok()) s = Check(cudaDeviceSynchronize(), "convert weight sync");
  cudaFree(staging);
  RET_IF_ERR(s);
  weight_bytes += static_cast<int64_t>(N) * ldk * 2;
  out->N = N; out->K = K; out->ldk = ldk; out->b = nullptr;
  if (bias_or_null != nullptr) RET_IF_ERR(UploadHostF32(bias_or_null, N, &out->b));
  return Status::OK();
}

Status Engine::LoadBlock(const OnnxModel& m, const std::string& p, bool timm, BlockW* b) {
  const std::string n1 = timm ? ".norm1" : ".ln_1", n2 = timm ? ".norm2" : ".ln_2";
  RET_IF_ERR(UploadF32(m, p + n1 + ".weight", D_, &b->ln1.g));
  RET_IF_ERR(UploadF32(m, p + n1 + ".bias", D_, &b->ln1.b));
  RET_IF_ERR(UploadF32(m, p + n2 + ".weight", D_, &b->ln2.g));
  RET_IF_ERR(UploadF32(m, p + n2 + ".bias", D_, &b->ln2.b));
  if (timm) {
    RET_IF_ERR(UploadLinear(m, p + ".attn.qkv.weight", p + ".attn.qkv.bias", 3 * D_, D_, false, &b->qkv));
    RET_IF_ERR(UploadLinear(m, p + ".attn.proj.weight", p + ".attn.proj.bias", D_, D_, false, &b->proj));
    RET_IF_ERR(UploadLinear(m, p + ".mlp.fc1.weight", p + ".mlp.fc1.bias", mlp_, D_, false, &b->fc1));
    RET_IF_ERR(UploadLinear(m, p + ".mlp.fc2.weight", p + ".mlp.fc2.bias", D_, mlp_, false, &b->fc2));
  } else {
    RET_IF_ERR(UploadLinear(m, p + ".attn.in_proj_weight", p + ".attn.in_proj_bias", 3 * D_, D_, false, &b->qkv));
    RET_IF_ERR(UploadLinear(m, p + ".attn.out_proj.weight", p + ".attn.out_proj.bias", D_, D_, false, &b->proj));
    RET_IF_ERR(UploadLinear(m, p + ".mlp.c_fc.weight", p + ".mlp.c_fc.bias", mlp_, D_, false, &b->fc1));
    RET_IF_ERR(UploadLinear(m, p + ".mlp.c_proj.weight", p + ".mlp.c_proj.bias", D_, mlp_, false, &b->fc2));
  }
  return Status::OK();
}

static int meta_int(const OnnxModel& m, const char* key, int dflt) {
  const std::string v = m.meta(std::string("clipb200.") + key);
  return v.empty() ? dflt : atoi(v.c_str());
}

Status Engine::LoadVision(const OnnxModel& m) {
  kind = CLIPB200_KIND_VISION;
  if (m.has("model.visual.trunk.stem.0.reparam_conv.weight")) return LoadFastVit(m);
  const bool timm = m.has("model.visual.trunk.patch_embed.proj.weight");
  const bool clip = m.has("model.visual.conv1.weight");
  if (!timm && !clip)
    return Status::Err(CLIPB200_ERR_UNSUPPORTED,
                       "vision graph has neither model.visual.conv1.weight (open_clip ViT) nor "
                       "model.visual.trunk.patch_embed.proj.weight (timm ViT); FastViT / renamed initializers are "
                       "not supported yet");
  family_ = timm ? "timm" : "clip";
  const std::string pre = timm ? "model.visual.trunk" : "model.visual";
  const OnnxTensor* pw = m.find(timm ? pre + ".patch_embed.proj.weight" : pre + ".conv1.weight");
  if (pw->dims.size() != 4 || pw->dims[1] != 3 || pw->dims[2] != pw->dims[3])
    return Status::Err(CLIPB200_ERR_UNSUPPORTED, "patch embedding weight must be [D,3,P,P]");
  D_ = static_cast<int>(pw->dims[0]);
  P_ = static_cast<int>(pw->dims[2]);
  const OnnxTensor* pos = m.find(timm ? pre + ".pos_embed" : pre + ".positional_embedding");
  if (pos == nullptr) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "positional embedding not found");
  T_ = static_cast<int>(pos->numel() / D_);
  has_cls_ = !timm;
  Tp_ = has_cls_ ? T_ - 1 : T_;
  G_ = static_cast<int>(lround(sqrt(static_cast<double>(Tp_))));
  if (G_ * G_ != Tp_) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "non-square patch grid");
  S_ = G_ * P_;
  image_size = S_;
  K_ = 3 * P_ * P_;
  Kp_ = (K_ + 7) & ~7;
  // depth: count blocks by probing names
  L_ = 0;
  while (m.has(pre + (timm ? ".blocks." : ".transformer.resblocks.") + std::to_string(L_) +
               (timm ? ".norm1.weight" : ".ln_1.weight")))
    ++L_;
  if (L_ == 0) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "no transformer blocks found");
  const OnnxTensor* fc1 = m.find(pre + (timm ? ".blocks.0.mlp.fc1.weight" : ".transformer.resblocks.0.mlp.c_fc.weight"));
  if (fc1 == nullptr) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "MLP weight not found");
  mlp_ = static_cast<int>(fc1->dims[0]);
  // heads are not derivable from shapes: metadata first, then the family defaults (head width 64; the
  // SigLIP so400m / giant-opt trunks use 16 heads, ViT-H/14 uses head width 80)
  H_ = meta_int(m, "heads", 0);
  if (H_ == 0) {
    if (D_ == 1152 || D_ == 1536 || D_ == 1280) H_ = 16;
    else H_ = D_ / 64;
  }
  if (H_ <= 0 || D_ % H_ != 0) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "bad head count");
  hd_ = D_ / H_;
  act_ = meta_int(m, "act", timm ? ACT_GELU_TANH : ACT_QUICKGELU);
  const std::string eps = m.meta("clipb200.eps");
  eps_ = eps.empty() ? (timm ? 1e-6f : 1e-5f) : static_cast<float>(atof(eps.c_str()));
  pool_map_ = timm;

  const OnnxTensor* out_w = timm ? nullptr : m.find(pre + ".proj");
  E_ = timm ? D_ : static_cast<int>(out_w ? out_w->dims.back() : 0);
  if (E_ <= 0) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "model.visual.proj not found");
  embed_dim = E_;

  RET_IF_ERR(UploadLinear(m, timm ? pre + ".patch_embed.proj.weight" : pre + ".conv1.weight",
                          timm ? pre + ".patch_embed.proj.bias" : "", D_, K_, false, &patch_));
  RET_IF_ERR(UploadF32(m, timm ? pre + ".pos_embed" : pre + ".positional_embedding", static_cast<int64_t>(T_) * D_, &pos_));
  if (has_cls_) {
    std::vector<float> cls, pos0;
    RET_IF_ERR(HostF32(m, pre + ".class_embedding", D_, &cls));
    RET_IF_ERR(HostF32(m, pre + ".positional_embedding", static_cast<int64_t>(T_) * D_, &pos0));
    for (int i = 0; i < D_; ++i) cls[i] += pos0[i];
    RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&cls_row_), D_ * 4));
    CUDA_RET(cudaMemcpy(cls_row_, cls.data(), D_ * 4, cudaMemcpyHostToDevice), "upload cls row");
    RET_IF_ERR(UploadF32(m, pre + ".ln_pre.weight", D_, &ln_pre_.g));
    RET_IF_ERR(UploadF32(m, pre + ".ln_pre.bias", D_, &ln_pre_.b));
  }
  blocks_.resize(L_);
  for (int i = 0; i < L_; ++i)
    RET_IF_ERR(LoadBlock(m, pre + (timm ? ".blocks." : ".transformer.resblocks.") + std::to_string(i), timm, &blocks_[i]));
  if (timm) {
    RET_IF_ERR(UploadF32(m, pre + ".norm.weight", D_, &ln_post_.g));
    RET_IF_ERR(UploadF32(m, pre + ".norm.bias", D_, &ln_post_.b));
    const std::string ap = pre + ".attn_pool";
    // the pooling query is weight-only: q = (latent @ Wq^T + bq) * hd^-0.5, folded at load time (the graph
    // recogniser folds it from the graph itself and hands it over as `clipb200.map_query`)
    std::vector<float> q(D_);
    if (m.has("clipb200.map_query")) {
      RET_IF_ERR(HostF32(m, "clipb200.map_query", D_, &q));
    } else {
      std::vector<float> latent, wq, bq;
      RET_IF_ERR(HostF32(m, ap + ".latent", D_, &latent));
      RET_IF_ERR(HostF32(m, ap + ".q.weight", static_cast<int64_t>(D_) * D_, &wq));
      RET_IF_ERR(HostF32(m, ap + ".q.bias", D_, &bq));
      const double sc = 1.0 / sqrt(static_cast<double>(hd_));
      for (int o = 0; o < D_; ++o) {
        double acc = bq[o];
        for (int i = 0; i < D_; ++i) acc += static_cast<double>(wq[static_cast<size_t>(o) * D_ + i]) * latent[i];
        q[o] = static_cast<float>(acc * sc);
      }
    }
    RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&map_q_), D_ * 4));
    CUDA_RET(cudaMemcpy(map_q_, q.data(), D_ * 4, cudaMemcpyHostToDevice), "upload pool query");
    RET_IF_ERR(UploadLinear(m, ap + ".kv.weight", ap + ".kv.bias", 2 * D_, D_, false, &map_kv_));
    RET_IF_ERR(UploadLinear(m, ap + ".proj.weight", ap + ".proj.bias", D_, D_, false, &map_proj_));
    RET_IF_ERR(UploadF32(m, ap + ".norm.weight", D_, &map_norm_.g));
    RET_IF_ERR(UploadF32(m, ap + ".norm.bias", D_, &map_norm_.b));
    RET_IF_ERR(UploadLinear(m, ap + ".mlp.fc1.weight", ap + ".mlp.fc1.bias", mlp_, D_, false, &map_fc1_));
    RET_IF_ERR(UploadLinear(m, ap + ".mlp.fc2.weight", ap + ".mlp.fc2.bias", D_, mlp_, false, &map_fc2_));
  } else {
    RET_IF_ERR(UploadF32(m, pre + ".ln_post.weight", D_, &ln_post_.g));
    RET_IF_ERR(UploadF32(m, pre + ".ln_post.bias", D_, &ln_post_.b));
    RET_IF_ERR(UploadLinear(m, pre + ".proj", "", E_, D_, true, &head_));
  }
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&lut_), 3 * 256 * 4));
  return Status::OK();
}

Status Engine::LoadText(const OnnxModel& m) {
  kind = CLIPB200_KIND_TEXT;
  std::string pre;
  if (m.has("model.token_embedding.weight")) pre = "model";
  else if (m.has("model.text.token_embedding.weight")) pre = "model.text";
  else
    return Status::Err(CLIPB200_ERR_UNSUPPORTED,
                       "text graph has no model.token_embedding.weight / model.text.token_embedding.weight");
  family_ = pre == "model" ? "clip" : "custom";
  const OnnxTensor* te = m.find(pre + ".token_embedding.weight");
  if (te->dims.size() != 2) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "token embedding must be 2-D");
  vocab_ = static_cast<int>(te->dims[0]);
  D_ = static_cast<int>(te->dims[1]);
  const OnnxTensor* pos = m.find(pre + ".positional_embedding");
  if (pos == nullptr) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "text positional embedding not found");
  T_ = static_cast<int>(pos->numel() / D_);
  context_length = T_;
  L_ = 0;
  while (m.has(pre + ".transformer.resblocks." + std::to_string(L_) + ".ln_1.weight")) ++L_;
  if (L_ == 0) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "no text transformer blocks found");
  const OnnxTensor* fc1 = m.find(pre + ".transformer.resblocks.0.mlp.c_fc.weight");
  if (fc1 == nullptr) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "text MLP weight not found");
  mlp_ = static_cast<int>(fc1->dims[0]);
  H_ = meta_int(m, "heads", 0);
  if (H_ == 0) H_ = (D_ == 1152) ? 16 : D_ / 64;
  if (H_ <= 0 || D_ % H_ != 0) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "bad text head count");
  hd_ = D_ / H_;
  const bool linear_head = m.has(pre + ".text_projection.weight");
  act_ = meta_int(m, "act", linear_head ? ACT_GELU_TANH : ACT_QUICKGELU);
  const std::string eps = m.meta("clipb200.eps");
  eps_ = eps.empty() ? (linear_head ? 1e-6f : 1e-5f) : static_cast<float>(atof(eps.c_str()));
  causal_ = meta_int(m, "causal", linear_head ? 0 : 1) != 0;
  const std::string pool = m.meta("clipb200.pool", linear_head ? "last" : "argmax");
  pool_argmax_ = pool == "argmax";
  if (linear_head) E_ = static_cast<int>(m.find(pre + ".text_projection.weight")->dims[0]);
  else {
    const OnnxTensor* tp = m.find(pre + ".text_projection");
    if (tp == nullptr) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "text_projection not found");
    E_ = static_cast<int>(tp->dims.back());
  }
  embed_dim = E_;
  RET_IF_ERR(UploadF32(m, pre + ".token_embedding.weight", static_cast<int64_t>(vocab_) * D_, &tok_emb_));
  RET_IF_ERR(UploadF32(m, pre + ".positional_embedding", static_cast<int64_t>(T_) * D_, &pos_));
  blocks_.resize(L_);
  for (int i = 0; i < L_; ++i)
    RET_IF_ERR(LoadBlock(m, pre + ".transformer.resblocks." + std::to_string(i), false, &blocks_[i]));
  RET_IF_ERR(UploadF32(m, pre + ".ln_final.weight", D_, &ln_post_.g));
  RET_IF_ERR(UploadF32(m, pre + ".ln_final.bias", D_, &ln_post_.b));
  if (linear_head)
    RET_IF_ERR(UploadLinear(m, pre + ".text_projection.weight", pre + ".text_projection.bias", E_, D_, false, &head_));
  else
    RET_IF_ERR(UploadLinear(m, pre + ".text_projection", "", E_, D_, true, &head_));
  return Status::OK();
}

// Activation buffers of ONE compute lane, allocated into the members Forward* use (SaveLane / BindLane move them).
Status Engine::AllocActivations() {
  const size_t rows = fastvit_ ? 1 : static_cast<size_t>(mb_) * T_;  // FastViT sizes its own buffers below
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&x_), rows * D_ * 4));
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&h_), rows * D_ * 2));
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&qkv_), rows * 3 * D_ * 2));
  if (attn_vt_)
    RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&vt_), static_cast<size_t>(mb_) * D_ * attn::attn_vt_ld(T_) * 2));
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&mlpbuf_), rows * mlp_ * 2));
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&pooled_), static_cast<size_t>(mb_) * D_ * 2));
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&proj_out_), static_cast<size_t>(mb_) * std::max(E_, D_) * 4));
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&row_map_), static_cast<size_t>(mb_) * 4));
  if (kind == CLIPB200_KIND_VISION && !fastvit_) {
    RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&patches_), static_cast<size_t>(mb_) * Tp_ * Kp_ * 2));
    if (pool_map_) {
      RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&y_), static_cast<size_t>(mb_) * D_ * 4));
      RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&yh_), static_cast<size_t>(mb_) * D_ * 2));
      RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&ymlp_), static_cast<size_t>(mb_) * mlp_ * 2));
    }
  }
  if (fastvit_) RET_IF_ERR(AllocFastVitWorkspace());
  return Status::OK();
}

void Engine::SaveLane(int k) {
  Lane& l = lanes_[k];
  l.patches = patches_; l.h = h_; l.qkv = qkv_; l.mlpbuf = mlpbuf_; l.pooled = pooled_; l.yh = yh_; l.ymlp = ymlp_;
  l.vt = vt_; l.fv_stem_out = fv_stem_out_; l.x = x_; l.y = y_; l.proj_out = proj_out_; l.fv_xa = fv_xa_; l.fv_xb = fv_xb_;
  l.fv_tmp = fv_tmp_; l.fv_s = fv_s_; l.fv_gate = fv_gate_; l.row_map = row_map_;
  l.stream = compute_;
}
void Engine::BindLane(int k) {
  if (k >= n_lanes_) k = 0;
  const Lane& l = lanes_[k];
  patches_ = l.patches; h_ = l.h; qkv_ = l.qkv; mlpbuf_ = l.mlpbuf; pooled_ = l.pooled; yh_ = l.yh; ymlp_ = l.ymlp;
  vt_ = l.vt; fv_stem_out_ = l.fv_stem_out; x_ = l.x; y_ = l.y; proj_out_ = l.proj_out; fv_xa_ = l.fv_xa; fv_xb_ = l.fv_xb;
  fv_tmp_ = l.fv_tmp; fv_s_ = l.fv_s; fv_gate_ = l.fv_gate; row_map_ = l.row_map;
  compute_ = l.stream;
  lane_ = k;
}

Status Engine::AllocWorkspace() {
  // Transposed V [mb][D][ld]: the qkv GEMM's epilogue writes V as [b][h*hd + d][t] and O += P V becomes one tcgen05.mma
  // per 16 keys.  Used where it measures faster (attn_vt_preferred: head dim 96 with T % 32 == 0, the giant-opt
  // SigLIP2 tower: +8 % on the attention kernel, nothing on the GEMM; at head dims 64 / 72 / 80 the natural layout is
  // as fast or faster once the item-boundary bubbles are gone, profiles/r02d_*).
  // CLIPB200_ATTN_VT=0 / 1 forces it off / on (wherever T % 32 == 0) for A/B runs.
  {
    const char* env = getenv("CLIPB200_ATTN_VT");
    const bool possible = !fastvit_ && attn_tcgen05_supported(hd_) && (2 * D_) % 32 == 0 && T_ % 32 == 0;
    attn_vt_ = possible && (env != nullptr ? atoi(env) != 0 : attn_vt_preferred(hd_, T_));
  }
  // CLIPB200_LANES=2 turns the second compute lane on (it costs one more activation workspace).  Off by default:
  // measured on B200 (profiles/r02h_*), two lanes overlap the HBM-bound kernels of one micro-batch with the GEMMs of the
  // other as intended, but the step is energy-bound under the 1 kW cap — the overlapped step draws more power, the clock
  // drops from 1.327 to 1.29 GHz and the throughput does not move (SO400M 1 938 -> 1 921 img/s, DFN5B text +1.4 %).
  n_lanes_ = 1;
  if (const char* env = getenv("CLIPB200_LANES")) n_lanes_ = atoi(env) >= 2 ? 2 : 1;
  for (int k = 0; k < n_lanes_; ++k) {
    compute_ = lanes_[k].stream;   // created in Init
    RET_IF_ERR(AllocActivations());
    SaveLane(k);
  }
  BindLane(0);
  CUDA_RET(cudaEventCreateWithFlags(&lane_fork_, cudaEventDisableTiming), "event");
  CUDA_RET(cudaEventCreateWithFlags(&lane_join_, cudaEventDisableTiming), "event");
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&err_flag_), 4));
  CUDA_RET(cudaMemset(err_flag_, 0, 4), "memset");
  if (kind == CLIPB200_KIND_VISION) in_slot_bytes_ = static_cast<size_t>(mb_) * S_ * S_ * 3;
  else in_slot_bytes_ = static_cast<size_t>(mb_) * T_ * 8;
  for (int i = 0; i < 2; ++i) {
    RET_IF_ERR(DevAlloc(&d_in_[i], in_slot_bytes_));
    RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&d_out_[i]), static_cast<size_t>(mb_) * E_ * 4));
    CUDA_RET(cudaHostAlloc(&h_in_[i], in_slot_bytes_, cudaHostAllocDefault), "pinned input staging");
    CUDA_RET(cudaHostAlloc(reinterpret_cast<void**>(&h_out_[i]), static_cast<size_t>(mb_) * E_ * 4, cudaHostAllocDefault),
             "pinned output staging");
    CUDA_RET(cudaEventCreateWithFlags(&in_ready_[i], cudaEventDisableTiming), "event");
    CUDA_RET(cudaEventCreateWithFlags(&in_consumed_[i], cudaEventDisableTiming), "event");
    CUDA_RET(cudaEventCreateWithFlags(&out_ready_[i], cudaEventDisableTiming), "event");
    CUDA_RET(cudaEventCreateWithFlags(&out_copied_[i], cudaEventDisableTiming), "event");
  }
  for (int i = 0; i < 16; ++i) CUDA_RET(cudaEventCreate(&user_events_[i]), "event");
  return Status::OK();
}

Status Engine::Init(const std::string& onnx_path, int dev, const clipb200_opts* opts) {
  device = dev;
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0)
    return Status::Err(CLIPB200_ERR_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") +
                                              cudaGetErrorString(ce));
  if (dev < 0 || dev >= count)
    return Status::Err(CLIPB200_ERR_INVALID_ARG, "cuda_device " + std::to_string(dev) + " out of range");
  CUDA_RET(cudaSetDevice(dev), "cudaSetDevice");
  cudaDeviceProp prop;
  CUDA_RET(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties");
  if (prop.major != 10)
    return Status::Err(CLIPB200_ERR_CUDA, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                              std::to_string(prop.minor) + "; this engine is built for sm_100a only");
  num_sms_ = prop.multiProcessorCount;

  OnnxModel m;
  std::string err;
  if (!load_onnx(onnx_path, &m, &err)) {
    const bool io = err.find("cannot open") != std::string::npos || err.find("cannot stat") != std::string::npos ||
                    err.find("cannot mmap") != std::string::npos || err.find("is empty") != std::string::npos;
    return Status::Err(io ? CLIPB200_ERR_IO : CLIPB200_ERR_PARSE, err);
  }
  input_names = m.inputs;
  // src/text.rs:156-161 feeds an attention_mask when the graph declares one.  pull_onnx.py's TextWrapper never does;
  // a graph that does uses the mask in a way only its nodes define, so it is refused instead of being ignored.
  for (const std::string& n : m.inputs)
    if (n == "attention_mask")
      return Status::Err(CLIPB200_ERR_UNSUPPORTED,
                         onnx_path + ": the graph declares an attention_mask input; this engine only runs text towers "
                                     "whose padding is handled inside the graph (pull_onnx.py:61-68)");
  // A file with an executable graph is bound ONLY from that graph: guessing hyper-parameters (activation, eps, heads,
  // causal mask, pooling) from parameter names would load such a file and return silently wrong embeddings where
  // onnxruntime executes what the graph says.  The one family bound by name is the re-parameterised FastViT trunk
  // (identified by its stem parameter), whose conv graph the recogniser does not parse.
  std::string graph_err;
  if (graph_needs_recognition(m) && !recognize_graph(&m, &graph_err, nullptr)) {
    if (!m.has("model.visual.trunk.stem.0.reparam_conv.weight"))
      return Status::Err(CLIPB200_ERR_UNSUPPORTED, onnx_path + ": " + graph_err);
    graph_note_ = graph_err;
    std::string fv_err;   // a real FastViT export: the attention blocks' Linear weights come out of the graph
    if (!bind_fastvit_graph(&m, &fv_err)) return Status::Err(CLIPB200_ERR_UNSUPPORTED, onnx_path + ": " + fv_err);
  }
  if (get_encode_tiled() == nullptr)  // resolved here so that it never happens inside a graph capture
    return Status::Err(CLIPB200_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
  if (const char* env = getenv("CLIPB200_NO_GRAPHS")) if (atoi(env) != 0) graph_max_n_ = 0;
  CUDA_RET(gemm_configure_device(), "configure GEMM kernels");
  CUDA_RET(flash_attention_configure_device(), "configure attention kernels");
  CUDA_RET(attn_configure_all(), "configure tcgen05 attention kernels");
  CUDA_RET(dwconv_tma_configure_device(), "configure depthwise-conv kernels");
  for (int k = 0; k < 2; ++k) CUDA_RET(cudaStreamCreateWithFlags(&lanes_[k].stream, cudaStreamNonBlocking), "stream");
  compute_ = lanes_[0].stream;
  {
    // copy-in stream (uploads) and resize stream (the photo path's resize kernels, so that the upload of staging group
    // g + 1 does not queue behind the resize of group g).  CLIPB200_COPY_STREAM_PRIORITY=1 creates both at the highest
    // priority; measured without effect on the photo workloads (profiles/r02z_photos.md), so the default is normal.
    int lo = 0, hi = 0;
    CUDA_RET(cudaDeviceGetStreamPriorityRange(&lo, &hi), "stream priority range");
    const char* pr = getenv("CLIPB200_COPY_STREAM_PRIORITY");
    const bool high = pr != nullptr && atoi(pr) != 0;
    CUDA_RET(cudaStreamCreateWithPriority(&copy_in_, cudaStreamNonBlocking, high ? hi : lo), "stream");
    CUDA_RET(cudaStreamCreateWithPriority(&resize_, cudaStreamNonBlocking, high ? hi : lo), "stream");
    if (const char* v = getenv("CLIPB200_PHOTO_STAGES")) rs_stages_ = std::min(std::max(atoi(v), 2), kResizeStages);
  }
  CUDA_RET(cudaStreamCreateWithFlags(&copy_out_, cudaStreamNonBlocking), "stream");

  bool is_text = false;
  for (const std::string& n : m.inputs) if (n == "input_ids") is_text = true;
  const std::string tower = m.meta("clipb200.tower");
  if (tower == "text") is_text = true;
  {
    Status ls = is_text ? LoadText(m) : LoadVision(m);
    if (!ls.ok() && !graph_note_.empty()) return Status::Err(ls.code, ls.msg + "; " + graph_note_);
    RET_IF_ERR(ls);
  }
  if (!fastvit_ && hd_ != 32 && hd_ != 64 && hd_ != 72 && hd_ != 80 && hd_ != 96 && hd_ != 128)
    return Status::Err(CLIPB200_ERR_UNSUPPORTED, "head_dim " + std::to_string(hd_) + " not supported");
  if (!fastvit_ && ((D_ & 7) || (mlp_ & 7) || (E_ & 7) || D_ > 2048))
    return Status::Err(CLIPB200_ERR_UNSUPPORTED, "width / mlp / embed_dim must be multiples of 8 and width <= 2048");

  profile_ = opts != nullptr && opts->profile != 0;
  mb_ = opts != nullptr ? opts->micro_batch : 0;
  if (const char* env = getenv("CLIPB200_MICRO_BATCH")) if (mb_ <= 0) mb_ = atoi(env);
  if (mb_ <= 0 && fastvit_) mb_ = 256;  // measured on B200: 64 -> 4.8k, 128 -> 5.3k, 256 -> 5.6k img/s (before the TMA dwconv)
  if (mb_ <= 0) {
    mb_ = 147456 / T_;  // ~147k token rows per step (256 SO400M images): measured best on B200 (32..1024 swept)
    if (mb_ > 1024) mb_ = 1024;
    if (mb_ < 1) mb_ = 1;
  }
  RET_IF_ERR(AllocWorkspace());
  CUDA_RET(cudaDeviceSynchronize(), "init sync");
  return Status::OK();
}

Engine::~Engine() {
  cudaSetDevice(device);
  cudaDeviceSynchronize();
  for (void* p : dev_allocs_) cudaFree(p);
  for (int i = 0; i < 2; ++i) {
    if (h_in_[i]) cudaFreeHost(h_in_[i]);
    if (h_in_f32_[i]) cudaFreeHost(h_in_f32_[i]);
    if (h_out_[i]) cudaFreeHost(h_out_[i]);
    if (d_in_f32_[i]) cudaFree(d_in_f32_[i]);
    if (in_ready_[i]) cudaEventDestroy(in_ready_[i]);
    if (in_consumed_[i]) cudaEventDestroy(in_consumed_[i]);
    if (out_ready_[i]) cudaEventDestroy(out_ready_[i]);
    if (out_copied_[i]) cudaEventDestroy(out_copied_[i]);
  }
  for (int i = 0; i < 16; ++i) if (user_events_[i]) cudaEventDestroy(user_events_[i]);
  for (auto& g : graphs_) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
  for (auto& p : prof_pending_) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (auto& e : prof_free_) cudaEventDestroy(e);
  if (l2_flush_) cudaFree(l2_flush_);
  for (ResizeStage& rs : rs_stage_) {
    if (rs.h_src) cudaFreeHost(rs.h_src);
    if (rs.h_arena) cudaFreeHost(rs.h_arena);
    if (rs.h_jobs) cudaFreeHost(rs.h_jobs);
    if (rs.d_src) cudaFree(rs.d_src);
    if (rs.d_tmp) cudaFree(rs.d_tmp);
    if (rs.d_arena) cudaFree(rs.d_arena);
    if (rs.d_jobs) cudaFree(rs.d_jobs);
    if (rs.free_ev) cudaEventDestroy(rs.free_ev);
    if (rs.h2d_ev) cudaEventDestroy(rs.h2d_ev);
  }
  for (int k = 0; k < 2; ++k) if (lanes_[k].stream) cudaStreamDestroy(lanes_[k].stream);
  if (lane_fork_) cudaEventDestroy(lane_fork_);
  if (lane_join_) cudaEventDestroy(lane_join_);
  if (copy_in_) cudaStreamDestroy(copy_in_);
  if (resize_) cudaStreamDestroy(resize_);
  if (copy_out_) cudaStreamDestroy(copy_out_);
}

// ------------------------------------------------------------------------------------------------ profiling
void Engine::ProfBegin(int cls, cudaStream_t st) {
  if (cls < PC_H2D || cls == PC_CONV) ++launch_count;  // kernels only; the two copy classes are DMA transfers
  if (!profile_) return;
  ProfPair p;
  p.cls = cls;
  auto get = [&]() {
    cudaEvent_t e;
    if (!prof_free_.empty()) { e = prof_free_.back(); prof_free_.pop_back(); }
    else cudaEventCreate(&e);
    return e;
  };
  p.a = get();
  p.b = get();
  cudaEventRecord(p.a, st);
  prof_pending_.push_back(p);
}
void Engine::ProfEnd(int cls, cudaStream_t st) {
  if (!profile_) return;
  (void)cls;
  cudaEventRecord(prof_pending_.back().b, st);
}

Status Engine::ReadProfile(clipb200_profile* out, bool reset) {
  CUDA_RET(cudaSetDevice(device), "cudaSetDevice");
  CUDA_RET(cudaDeviceSynchronize(), "profile sync");
  for (auto& p : prof_pending_) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      prof_acc_.ms[p.cls] += ms;
      prof_acc_.launches[p.cls] += 1;
    }
    prof_free_.push_back(p.a);
    prof_free_.push_back(p.b);
  }
  prof_pending_.clear();
  if (out != nullptr) *out = prof_acc_;
  if (reset) memset(&prof_acc_, 0, sizeof(prof_acc_));
  return Status::OK();
}

Status Engine::RecordEvent(int slot) {
  if (slot < 0 || slot >= 16) return Status::Err(CLIPB200_ERR_INVALID_ARG, "event slot out of range");
  CUDA_RET(cudaSetDevice(device), "cudaSetDevice");
  CUDA_RET(cudaEventRecord(user_events_[slot], compute_), "cudaEventRecord");
  return Status::OK();
}
Status Engine::ElapsedMs(int a, int b, double* ms) {
  if (a < 0 || a >= 16 || b < 0 || b >= 16 || ms == nullptr)
    return Status::Err(CLIPB200_ERR_INVALID_ARG, "event slot out of range");
  CUDA_RET(cudaSetDevice(device), "cudaSetDevice");
  CUDA_RET(cudaEventSynchronize(user_events_[b]), "cudaEventSynchronize");
  float f = 0.f;
  CUDA_RET(cudaEventElapsedTime(&f, user_events_[a], user_events_[b]), "cudaEventElapsedTime");
  *ms = f;
  return Status::OK();
}
Status Engine::Synchronize() {
  CUDA_RET(cudaSetDevice(device), "cudaSetDevice");
  CUDA_RET(cudaStreamSynchronize(copy_in_), "sync");
  CUDA_RET(cudaStreamSynchronize(resize_), "sync");
  for (int k = 0; k < n_lanes_; ++k) CUDA_RET(cudaStreamSynchronize(lanes_[k].stream), "sync");
  CUDA_RET(cudaStreamSynchronize(copy_out_), "sync");
  return Status::OK();
}
Status Engine::FlushL2() {
  CUDA_RET(cudaSetDevice(device), "cudaSetDevice");
  const size_t bytes = 256u << 20;  // 2x the 126 MB L2
  if (l2_flush_ == nullptr) CUDA_RET(cudaMalloc(&l2_flush_, bytes), "l2 flush buffer");
  CUDA_RET(cudaMemsetAsync(l2_flush_, 0x5a, bytes, compute_), "l2 flush");
  return Status::OK();
}

// ------------------------------------------------------------------------------------------------ forward
Status Engine::Gemm(const __nv_bfloat16* A, long long lda, const LinearW& w, int M, int epi, GemmEpilogue* ep) {
  if (ep->bias == nullptr) ep->bias = w.b;
  ProfBegin(PC_GEMM, compute_);
  cudaError_t e = gemm_bf16(A, lda, w.w, w.ldk, M, w.N, w.ldk, epi, *ep, num_sms_, compute_);
  ProfEnd(PC_GEMM, compute_);
  if (profile_) prof_acc_.gemm_flops += 2.0 * M * static_cast<double>(w.N) * w.K;
  return Check(e, "gemm launch");
}

Status Engine::Blocks(int rows, int n_seq, int T, bool causal) {
  for (int l = 0; l < L_; ++l) {
    const BlockW& b = blocks_[l];
    ProfBegin(PC_LN, compute_);
    cudaError_t e = launch_layernorm(x_, nullptr, rows, D_, b.ln1.g, b.ln1.b, eps_, h_, nullptr, compute_);
    ProfEnd(PC_LN, compute_);
    CUDA_RET(e, "layernorm");
    GemmEpilogue ep;
    ep.out_bf16 = qkv_;
    ep.ldc = 3 * D_;
    // With the transposed-V attention kernel the qkv GEMM writes q | k as before and v straight into
    // vt_[b][h*hd + d][t] (EPI_QKVT): V is then a K-major operand and O += P V is one tcgen05.mma per 16 keys.
    const bool vt = attn_vt_ && vt_ != nullptr;
    if (vt) {
      ep.out_vt = vt_;
      ep.vt_col0 = 2 * D_;
      ep.vt_T = T;
      ep.vt_ld = attn::attn_vt_ld(T);
      ep.vt_B = n_seq;
    }
    RET_IF_ERR(Gemm(h_, D_, b.qkv, rows, vt ? EPI_QKVT : EPI_BF16, &ep));
    ProfBegin(PC_ATTN, compute_);
    // tcgen05/TMEM kernel for head dims >= 64; the mma.sync kernel remains for head dim 32 (FastViT-style heads)
    if (vt) e = attn_tcgen05_vt(qkv_, vt_, h_, n_seq, T, H_, hd_, causal, num_sms_, compute_);
    else
      e = attn_tcgen05_supported(hd_) ? attn_auto(qkv_, h_, n_seq, T, H_, hd_, causal, num_sms_, compute_)
                                      : launch_flash_attention(qkv_, h_, n_seq, T, H_, hd_, causal, compute_);
    ProfEnd(PC_ATTN, compute_);
    CUDA_RET(e, "attention");
    GemmEpilogue ep2;
    ep2.out_f32 = x_;
    ep2.ldc = D_;
    RET_IF_ERR(Gemm(h_, D_, b.proj, rows, EPI_RESID, &ep2));
    ProfBegin(PC_LN, compute_);
    e = launch_layernorm(x_, nullptr, rows, D_, b.ln2.g, b.ln2.b, eps_, h_, nullptr, compute_);
    ProfEnd(PC_LN, compute_);
    CUDA_RET(e, "layernorm");
    GemmEpilogue ep3;
    ep3.out_bf16 = mlpbuf_;
    ep3.ldc = mlp_;
    ep3.act = act_;
    RET_IF_ERR(Gemm(h_, D_, b.fc1, rows, EPI_BF16, &ep3));
    GemmEpilogue ep4;
    ep4.out_f32 = x_;
    ep4.ldc = D_;
    RET_IF_ERR(Gemm(mlpbuf_, mlp_, b.fc2, rows, EPI_RESID, &ep4));
  }
  return Status::OK();
}

Status Engine::SetPreproc(const clipb200_preproc* pp) {
  if (pp == nullptr) return Status::Err(CLIPB200_ERR_INVALID_ARG, "preproc config is null");
  if (lut_valid_ && memcmp(lut_mean_, pp->mean, 12) == 0 && memcmp(lut_std_, pp->std, 12) == 0) return Status::OK();
  // vision.rs:253-254, evaluated in fp32 exactly as written there, for each of the 256 byte values
  float lut[3 * 256];
  for (int c = 0; c < 3; ++c)
    for (int v = 0; v < 256; ++v) {
      volatile float val = static_cast<float>(v) / 255.0f;
      volatile float d = val - pp->mean[c];
      lut[c * 256 + v] = d / pp->std[c];
    }
  CUDA_RET(cudaMemcpyAsync(lut_, lut, sizeof(lut), cudaMemcpyHostToDevice, compute_), "upload LUT");
  CUDA_RET(cudaStreamSynchronize(compute_), "upload LUT sync");
  memcpy(lut_mean_, pp->mean, 12);
  memcpy(lut_std_, pp->std, 12);
  lut_valid_ = true;
  return Status::OK();
}

Status Engine::ForwardVision(int n, const uint8_t* d_u8, const float* d_f32, float* d_out) {
  if (fastvit_) return ForwardFastVit(n, d_u8, d_f32, d_out);
  const int rows = n * T_;
  ProfBegin(PC_PRE, compute_);
  cudaError_t e = d_u8 != nullptr ? launch_preprocess_patches_u8(d_u8, n, S_, P_, Kp_, lut_, patches_, compute_)
                                  : launch_im2col_f32(d_f32, n, S_, P_, Kp_, patches_, compute_);
  ProfEnd(PC_PRE, compute_);
  CUDA_RET(e, "preprocess");
  {
    GemmEpilogue ep;
    ep.out_f32 = x_;
    ep.ldc = D_;
    ep.pos = pos_;
    ep.rows_in = Tp_;
    ep.rows_out = T_;
    ep.row_off = has_cls_ ? 1 : 0;
    RET_IF_ERR(Gemm(patches_, Kp_, patch_, n * Tp_, EPI_F32, &ep));
  }
  if (has_cls_) {
    ProfBegin(PC_MISC, compute_);
    e = launch_write_cls_rows(x_, n, T_, D_, cls_row_, compute_);
    ProfEnd(PC_MISC, compute_);
    CUDA_RET(e, "cls rows");
    ProfBegin(PC_LN, compute_);
    e = launch_layernorm(x_, nullptr, rows, D_, ln_pre_.g, ln_pre_.b, eps_, nullptr, x_, compute_);
    ProfEnd(PC_LN, compute_);
    CUDA_RET(e, "ln_pre");
  }
  RET_IF_ERR(Blocks(rows, n, T_, false));
  if (!pool_map_) {
    // open_clip ViT: ln_post on the class token, then x @ proj
    ProfBegin(PC_MISC, compute_);
    e = launch_affine_rows(n, T_, 0, row_map_, compute_);
    ProfEnd(PC_MISC, compute_);
    CUDA_RET(e, "row map");
    ProfBegin(PC_LN, compute_);
    e = launch_layernorm(x_, row_map_, n, D_, ln_post_.g, ln_post_.b, eps_, pooled_, nullptr, compute_);
    ProfEnd(PC_LN, compute_);
    CUDA_RET(e, "ln_post");
    GemmEpilogue ep;
    ep.out_f32 = proj_out_;
    ep.ldc = E_;
    RET_IF_ERR(Gemm(pooled_, D_, head_, n, EPI_F32, &ep));
    ProfBegin(PC_MISC, compute_);
    e = launch_l2_normalize(proj_out_, n, E_, d_out, compute_);
    ProfEnd(PC_MISC, compute_);
    CUDA_RET(e, "l2 normalize");
  } else {
    // timm: final norm, then AttentionPoolLatent (one latent query), then x + mlp(norm(x))
    ProfBegin(PC_LN, compute_);
    e = launch_layernorm(x_, nullptr, rows, D_, ln_post_.g, ln_post_.b, eps_, h_, nullptr, compute_);
    ProfEnd(PC_LN, compute_);
    CUDA_RET(e, "final norm");
    GemmEpilogue ep;
    ep.out_bf16 = qkv_;
    ep.ldc = 2 * D_;
    RET_IF_ERR(Gemm(h_, D_, map_kv_, rows, EPI_BF16, &ep));
    ProfBegin(PC_ATTN, compute_);
    e = launch_map_pool_attention(qkv_, map_q_, pooled_, n, T_, H_, hd_, compute_);
    ProfEnd(PC_ATTN, compute_);
    CUDA_RET(e, "attention pool");
    GemmEpilogue ep2;
    ep2.out_f32 = y_;
    ep2.ldc = D_;
    RET_IF_ERR(Gemm(pooled_, D_, map_proj_, n, EPI_F32, &ep2));
    ProfBegin(PC_LN, compute_);
    e = launch_layernorm(y_, nullptr, n, D_, map_norm_.g, map_norm_.b, eps_, yh_, nullptr, compute_);
    ProfEnd(PC_LN, compute_);
    CUDA_RET(e, "pool norm");
    GemmEpilogue ep3;
    ep3.out_bf16 = ymlp_;
    ep3.ldc = mlp_;
    ep3.act = act_;
    RET_IF_ERR(Gemm(yh_, D_, map_fc1_, n, EPI_BF16, &ep3));
    GemmEpilogue ep4;
    ep4.out_f32 = y_;
    ep4.ldc = D_;
    RET_IF_ERR(Gemm(ymlp_, mlp_, map_fc2_, n, EPI_RESID, &ep4));
    ProfBegin(PC_MISC, compute_);
    e = launch_l2_normalize(y_, n, D_, d_out, compute_);
    ProfEnd(PC_MISC, compute_);
    CUDA_RET(e, "l2 normalize");
  }
  return Status::OK();
}

Status Engine::ForwardText(int n, const int64_t* d_ids, float* d_out) {
  const int rows = n * T_;
  ProfBegin(PC_MISC, compute_);
  cudaError_t e = launch_embed_tokens(d_ids, rows, T_, D_, vocab_, tok_emb_, pos_, x_, err_flag_, compute_);
  ProfEnd(PC_MISC, compute_);
  CUDA_RET(e, "token embedding");
  RET_IF_ERR(Blocks(rows, n, T_, causal_));
  ProfBegin(PC_MISC, compute_);
  e = launch_text_pool_rows(d_ids, n, T_, pool_argmax_, row_map_, compute_);
  ProfEnd(PC_MISC, compute_);
  CUDA_RET(e, "text pool rows");
  ProfBegin(PC_LN, compute_);
  e = launch_layernorm(x_, row_map_, n, D_, ln_post_.g, ln_post_.b, eps_, pooled_, nullptr, compute_);
  ProfEnd(PC_LN, compute_);
  CUDA_RET(e, "ln_final");
  GemmEpilogue ep;
  ep.out_f32 = proj_out_;
  ep.ldc = E_;
  RET_IF_ERR(Gemm(pooled_, D_, head_, n, EPI_F32, &ep));
  ProfBegin(PC_MISC, compute_);
  e = launch_l2_normalize(proj_out_, n, E_, d_out, compute_);
  ProfEnd(PC_MISC, compute_);
  CUDA_RET(e, "l2 normalize");
  return Status::OK();
}

// Forward pass of one micro-batch from staging slot `slot`.  Small micro-batches are launch-bound (a ViT-B/32 image
// is ~90 kernels of a few microseconds each), so their kernel sequence is captured once into a CUDA graph per
// (mode, n, slot) and replayed; tensor maps are by-value kernel parameters, so the captured launches are complete.
Status Engine::ForwardSlot(int mode, int n, int slot) {
  BindLane(slot);   // slot s & 1 <-> compute lane s & 1: the caller's event waits / records below use `compute_` too
  auto fwd = [&]() -> Status {
    if (mode == 0) return ForwardVision(n, static_cast<const uint8_t*>(d_in_[slot]), nullptr, d_out_[slot]);
    if (mode == 1) return ForwardVision(n, nullptr, d_in_f32_[slot], d_out_[slot]);
    return ForwardText(n, static_cast<const int64_t*>(d_in_[slot]), d_out_[slot]);
  };
  if (profile_ || n > graph_max_n_) return fwd();
  const auto key = std::make_tuple(mode, n, slot);
  auto it = graphs_.find(key);
  if (it == graphs_.end()) {
    const int64_t before = launch_count;
    CUDA_RET(cudaStreamBeginCapture(compute_, cudaStreamCaptureModeThreadLocal), "begin graph capture");
    Status st = fwd();
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(compute_, &graph);
    if (!st.ok()) {
      if (graph) cudaGraphDestroy(graph);
      return st;
    }
    CUDA_RET(ce, "end graph capture");
    GraphEntry g;
    g.launches = launch_count - before;
    launch_count = before;
    ce = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    CUDA_RET(ce, "instantiate graph");
    it = graphs_.emplace(key, g).first;
  }
  launch_count += it->second.launches;
  return Check(cudaGraphLaunch(it->second.exec, compute_), "graph launch");
}

// ------------------------------------------------------------------------------------------------ pipeline
// mode 0: vision u8, 1: vision f32 NCHW, 2: text ids
template <typename InT>
Status Engine::RunPipelined(const InT* in, int64_t batch, size_t in_elems_per_item, float* out, bool device_buffers,
                            int mode) {
  CUDA_RET(cudaSetDevice(device), "cudaSetDevice");
  const size_t item_bytes = in_elems_per_item * sizeof(InT);
  const int64_t steps = (batch + mb_ - 1) / mb_;
  if (device_buffers) {
    // micro-batches alternate between the compute lanes; lane 1 starts behind whatever the caller queued on lane 0
    // (its events, earlier calls) and lane 0 ends behind lane 1, so the call stays "asynchronous on ONE stream" for
    // the caller's event records
    BindLane(0);
    if (n_lanes_ > 1 && steps > 1) {
      CUDA_RET(cudaEventRecord(lane_fork_, lanes_[0].stream), "record");
      CUDA_RET(cudaStreamWaitEvent(lanes_[1].stream, lane_fork_, 0), "fork lane");
    }
    Status st = Status::OK();
    for (int64_t s = 0; s < steps && st.ok(); ++s) {
      const int n = static_cast<int>(std::min<int64_t>(mb_, batch - s * mb_));
      const InT* src = in + static_cast<size_t>(s) * mb_ * in_elems_per_item;
      float* dst = out + static_cast<size_t>(s) * mb_ * E_;
      BindLane(static_cast<int>(s & 1));
      if (mode == 0) st = ForwardVision(n, reinterpret_cast<const uint8_t*>(src), nullptr, dst);
      else if (mode == 1) st = ForwardVision(n, nullptr, reinterpret_cast<const float*>(src), dst);
      else st = ForwardText(n, reinterpret_cast<const int64_t*>(src), dst);
    }
    if (n_lanes_ > 1 && steps > 1) {
      cudaEventRecord(lane_join_, lanes_[1].stream);
      cudaStreamWaitEvent(lanes_[0].stream, lane_join_, 0);
    }
    BindLane(0);
    return st;  // asynchronous on compute lane 0; caller synchronises / records events
  }
  // host buffers: pinned staging unless the caller's memory is already pinned
  cudaPointerAttributes attr;
  bool in_pinned = cudaPointerGetAttributes(&attr, in) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  bool out_pinned = cudaPointerGetAttributes(&attr, out) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  void* const* d_in = mode == 1 ? reinterpret_cast<void* const*>(d_in_f32_) : d_in_;
  void* const* h_in = mode == 1 ? reinterpret_cast<void* const*>(h_in_f32_) : h_in_;
  Status st = Status::OK();
  for (int64_t s = 0; s < steps && st.ok(); ++s) {
    const int slot = static_cast<int>(s & 1);
    const int n = static_cast<int>(std::min<int64_t>(mb_, batch - s * mb_));
    const InT* src = in + static_cast<size_t>(s) * mb_ * in_elems_per_item;
    const size_t bytes = static_cast<size_t>(n) * item_bytes;
    if (s >= 2) {
      // slot reuse: the H2D of step s-2 must have left the pinned buffer, its kernels must have consumed d_in
      CUDA_RET(cudaEventSynchronize(in_ready_[slot]), "wait staging");
      CUDA_RET(cudaStreamWaitEvent(copy_in_, in_consumed_[slot], 0), "wait consumed");
    }
    const void* hsrc = src;
    if (!in_pinned) {
      memcpy(h_in[slot], src, bytes);
      hsrc = h_in[slot];
    }
    ProfBegin(PC_H2D, copy_in_);
    cudaError_t e = cudaMemcpyAsync(d_in[slot], hsrc, bytes, cudaMemcpyHostToDevice, copy_in_);
    ProfEnd(PC_H2D, copy_in_);
    CUDA_RET(e, "H2D copy");
    CUDA_RET(cudaEventRecord(in_ready_[slot], copy_in_), "record");
    BindLane(slot);
    CUDA_RET(cudaStreamWaitEvent(compute_, in_ready_[slot], 0), "wait input");
    if (s >= 2) CUDA_RET(cudaStreamWaitEvent(compute_, out_copied_[slot], 0), "wait output slot");
    st = ForwardSlot(mode, n, slot);
    if (!st.ok()) break;
    CUDA_RET(cudaEventRecord(in_consumed_[slot], compute_), "record");
    CUDA_RET(cudaEventRecord(out_ready_[slot], compute_), "record");
    CUDA_RET(cudaStreamWaitEvent(copy_out_, out_ready_[slot], 0), "wait output");
    float* user_dst = out + static_cast<size_t>(s) * mb_ * E_;
    ProfBegin(PC_D2H, copy_out_);
    e = cudaMemcpyAsync(out_pinned ? user_dst : h_out_[slot], d_out_[slot], static_cast<size_t>(n) * E_ * 4,
                        cudaMemcpyDeviceToHost, copy_out_);
    ProfEnd(PC_D2H, copy_out_);
    CUDA_RET(e, "D2H copy");
    CUDA_RET(cudaEventRecord(out_copied_[slot], copy_out_), "record");
    // drain the previous step's output while this step runs
    if (!out_pinned && s >= 1) {
      const int ps = static_cast<int>((s - 1) & 1);
      const int pn = static_cast<int>(std::min<int64_t>(mb_, batch - (s - 1) * mb_));
      CUDA_RET(cudaEventSynchronize(out_copied_[ps]), "wait D2H");
      memcpy(out + static_cast<size_t>(s - 1) * mb_ * E_, h_out_[ps], static_cast<size_t>(pn) * E_ * 4);
    }
  }
  BindLane(0);
  Status sync = Synchronize();
  if (!st.ok()) return st;
  RET_IF_ERR(sync);
  if (!out_pinned) {
    const int64_t s = steps - 1;
    const int n = static_cast<int>(std::min<int64_t>(mb_, batch - s * mb_));
    memcpy(out + static_cast<size_t>(s) * mb_ * E_, h_out_[s & 1], static_cast<size_t>(n) * E_ * 4);
  }
  if (mode == 2) {
    int flag = 0;
    CUDA_RET(cudaMemcpy(&flag, err_flag_, 4, cudaMemcpyDeviceToHost), "read error flag");
    if (flag != 0) {
      cudaMemset(err_flag_, 0, 4);
      return Status::Err(CLIPB200_ERR_INVALID_ARG, "input_ids contains an id outside [0, vocab_size)");
    }
  }
  return Status::OK();
}

Status Engine::VisionEmbedRgb8(const uint8_t* hwc, int64_t batch, int width, int height, const clipb200_preproc* pp,
                               float* out, bool device_buffers) {
  if (kind != CLIPB200_KIND_VISION) return Status::Err(CLIPB200_ERR_INVALID_ARG, "not a vision engine");
  if (batch <= 0) return Status::Err(CLIPB200_ERR_INVALID_ARG, "Empty batch");
  if (hwc == nullptr || out == nullptr) return Status::Err(CLIPB200_ERR_INVALID_ARG, "null buffer");
  if (width != S_ || height != S_)
    return Status::Err(CLIPB200_ERR_UNSUPPORTED, "images must already be " + std::to_string(S_) + "x" + std::to_string(S_) +
                                                     "; use clipb200_vision_embed_rgb8_var / clipb200_resize_rgb8 for other sizes");
  CUDA_RET(cudaSetDevice(device), "cudaSetDevice");
  RET_IF_ERR(SetPreproc(pp));
  return RunPipelined<uint8_t>(hwc, batch, static_cast<size_t>(S_) * S_ * 3, out, device_buffers, 0);
}

// The reference's public `preprocess_batch` (vision.rs:120-135): u8 HWC -> normalised f32 NCHW, on the GPU.
Status Engine::PreprocessRgb8(const uint8_t* hwc, int64_t batch, int width, int height, const clipb200_preproc* pp,
                              float* out_nchw) {
  if (kind != CLIPB200_KIND_VISION) return Status::Err(CLIPB200_ERR_INVALID_ARG, "not a vision engine");
  if (batch <= 0) return Status::Err(CLIPB200_ERR_INVALID_ARG, "Empty batch");
  if (hwc == nullptr || out_nchw == nullptr) return Status::Err(CLIPB200_ERR_INVALID_ARG, "null buffer");
  if (width != S_ || height != S_)
    return Status::Err(CLIPB200_ERR_UNSUPPORTED, "images must already be " + std::to_string(S_) + "x" + std::to_string(S_) +
                                                     "; use clipb200_vision_embed_rgb8_var / clipb200_resize_rgb8 for other sizes");
  CUDA_RET(cudaSetDevice(device), "cudaSetDevice");
  RET_IF_ERR(SetPreproc(pp));
  const size_t px = static_cast<size_t>(S_) * S_ * 3;
  const int64_t chunk = 256;
  uint8_t* d_u8 = nullptr;
  float* d_f = nullptr;
  CUDA_RET(cudaMalloc(reinterpret_cast<void**>(&d_u8), chunk * px), "preprocess alloc");
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&d_f), chunk * px * 4);
  for (int64_t i = 0; i < batch && e == cudaSuccess; i += chunk) {
    const int n = static_cast<int>(std::min<int64_t>(chunk, batch - i));
    e = cudaMemcpyAsync(d_u8, hwc + i * px, n * px, cudaMemcpyHostToDevice, compute_);
    if (e == cudaSuccess) {
      ProfBegin(PC_PRE, compute_);
      e = launch_normalize_nchw_f32(d_u8, n, S_, lut_, d_f, compute_);
      ProfEnd(PC_PRE, compute_);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_nchw + i * px, d_f, n * px * 4, cudaMemcpyDeviceToHost, compute_);
    if (e == cudaSuccess) e = cudaStreamSynchronize(compute_);
  }
  cudaFree(d_u8);
  cudaFree(d_f);
  return Check(e, "preprocess");
}

// ------------------------------------------------------------------------------------------------ arbitrary sizes
const Engine::AxisEntry& Engine::GetAxis(int in_size, double in0, double in1, int interpolation) {
  AxisKey key;
  key.in_size = in_size;
  key.interp = interpolation;
  memcpy(&key.in0_bits, &in0, 8);
  memcpy(&key.in1_bits, &in1, 8);
  auto it = axis_cache_.find(key);
  if (it == axis_cache_.end()) {
    // bounded LRU: a long-running service sees arbitrarily many photo sizes
    constexpr size_t kMaxEntries = 512, kMaxBytes = size_t(64) << 20;
    while (!axis_cache_.empty() && (axis_cache_.size() >= kMaxEntries || axis_cache_bytes_ > kMaxBytes)) {
      auto victim = axis_cache_.begin();
      for (auto i = axis_cache_.begin(); i != axis_cache_.end(); ++i)
        if (i->second.tick < victim->second.tick) victim = i;
      axis_cache_bytes_ -= victim->second.bytes;
      axis_cache_.erase(victim);
    }
    AxisEntry e;
    e.axis = make_resize_axis(in_size, in0, in1, S_, interpolation);
    e.first = in_size;
    e.last = 0;
    for (int o = 0; o < S_; ++o) {
      e.first = std::min(e.first, e.axis.start[o]);
      e.last = std::max(e.last, e.axis.start[o] + e.axis.size[o]);
    }
    e.bytes = e.axis.w.size() * 2 + static_cast<size_t>(S_) * 8 + sizeof(AxisEntry);
    axis_cache_bytes_ += e.bytes;
    it = axis_cache_.emplace(key, std::move(e)).first;
  }
  it->second.tick = ++axis_tick_;
  return it->second;
}

Status Engine::GrowStage(ResizeStage* st, size_t src, size_t tmp, size_t arena_words, size_t jobs) {
  if (st->free_ev == nullptr) CUDA_RET(cudaEventCreateWithFlags(&st->free_ev, cudaEventDisableTiming), "event");
  if (st->h2d_ev == nullptr) CUDA_RET(cudaEventCreateWithFlags(&st->h2d_ev, cudaEventDisableTiming), "event");
  const bool grow = src > st->src_cap || tmp > st->tmp_cap || arena_words > st->arena_cap || jobs > st->jobs_cap;
  if (!grow) return Status::OK();
  CUDA_RET(cudaStreamSynchronize(copy_in_), "sync before growing the resize staging");
  auto round_up = [](size_t v, size_t q) { return (v + q - 1) / q * q; };
  if (src > st->src_cap) {
    if (st->h_src) cudaFreeHost(st->h_src);
    if (st->d_src) cudaFree(st->d_src);
    st->h_src = nullptr; st->d_src = nullptr; st->src_cap = 0;
    const size_t cap = round_up(src, size_t(16) << 20);
    CUDA_RET(cudaHostAlloc(reinterpret_cast<void**>(&st->h_src), cap, cudaHostAllocDefault), "pinned resize source staging");
    CUDA_RET(cudaMalloc(reinterpret_cast<void**>(&st->d_src), cap), "resize source buffer");
    st->src_cap = cap;
  }
  if (tmp > st->tmp_cap) {
    if (st->d_tmp) cudaFree(st->d_tmp);
    st->d_tmp = nullptr; st->tmp_cap = 0;
    const size_t cap = round_up(tmp, size_t(4) << 20);
    CUDA_RET(cudaMalloc(reinterpret_cast<void**>(&st->d_tmp), cap), "resize intermediate buffer");
    st->tmp_cap = cap;
  }
  if (arena_words > st->arena_cap) {
    if (st->h_arena) cudaFreeHost(st->h_arena);
    if (st->d_arena) cudaFree(st->d_arena);
    st->h_arena = nullptr; st->d_arena = nullptr; st->arena_cap = 0;
    const size_t cap = round_up(arena_words, size_t(1) << 18);
    CUDA_RET(cudaHostAlloc(reinterpret_cast<void**>(&st->h_arena), cap * 4, cudaHostAllocDefault), "pinned coefficient arena");
    CUDA_RET(cudaMalloc(reinterpret_cast<void**>(&st->d_arena), cap * 4), "coefficient arena");
    st->arena_cap = cap;
  }
  if (jobs > st->jobs_cap) {
    if (st->h_jobs) cudaFreeHost(st->h_jobs);
    if (st->d_jobs) cudaFree(st->d_jobs);
    st->h_jobs = nullptr; st->d_jobs = nullptr; st->jobs_cap = 0;
    const size_t cap = round_up(jobs, 256);
    CUDA_RET(cudaHostAlloc(reinterpret_cast<void**>(&st->h_jobs), cap * sizeof(ResizeJob), cudaHostAllocDefault), "pinned resize jobs");
    CUDA_RET(cudaMalloc(reinterpret_cast<void**>(&st->d_jobs), cap * sizeof(ResizeJob)), "resize jobs");
    st->jobs_cap = cap;
  }
  return Status::OK();
}

// Copies a list of (dst, src, bytes) with up to 16 host threads: one core moves ~10 GB/s from pageable memory, a photo
// is tens of MB, and PCIe takes > 50 GB/s, so a single-threaded staging copy would be the bottleneck of the whole path.
namespace {
struct CopyPiece {   // `rows` runs of `bytes` each, `src_pitch` apart in the source, packed in the destination
  uint8_t* dst;
  const uint8_t* src;
  size_t bytes;
  size_t rows = 1, src_pitch = 0;
};
// Staging copies are issued group by group (a few dozen per call): spawning 15 threads for each cost the staging host
// thread ~0.4 ms a time.  One process-wide set of workers (all engines, all pool replicas) sleeps on a condition variable
// and pulls 4 MB chunks from whichever jobs are posted; the posting thread works on its own job too and returns when
// that job's chunks are done.
struct CopyJob {
  const std::vector<CopyPiece>* chunks = nullptr;
  void (*run)(const CopyPiece&) = nullptr;
  std::atomic<size_t> next{0}, done{0};
  int workers = 0;   // pool threads currently holding a pointer to this job (guarded by the pool mutex)
};
class CopyWorkers {
 public:
  static CopyWorkers& get() {
    static CopyWorkers* w = new CopyWorkers();   // never destroyed: workers may outlive static destruction order
    return *w;
  }
  size_t size() const { return threads_.size(); }
  void run(CopyJob* job) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      jobs_.push_back(job);
    }
    cv_.notify_all();
    work_on(job);
    {   // every chunk has been claimed; wait until the workers that claimed the last ones have finished them
      std::unique_lock<std::mutex> lk(mu_);
      jobs_.erase(std::find(jobs_.begin(), jobs_.end(), job));
      done_cv_.wait(lk, [&] { return job->done.load() == job->chunks->size() && job->workers == 0; });
    }
  }

 private:
  CopyWorkers() {
    const unsigned hw = std::thread::hardware_concurrency();
    const size_t n = std::min<size_t>(15, hw > 2 ? hw - 2 : 0);
    for (size_t i = 0; i < n; ++i) threads_.emplace_back([this] { loop(); });
    for (std::thread& t : threads_) t.detach();
  }
  void work_on(CopyJob* job) {
    for (;;) {
      const size_t i = job->next.fetch_add(1);
      if (i >= job->chunks->size()) return;
      job->run((*job->chunks)[i]);
      if (job->done.fetch_add(1) + 1 == job->chunks->size()) {
        std::lock_guard<std::mutex> lk(mu_);
        done_cv_.notify_all();
      }
    }
  }
  void loop() {
    for (;;) {
      CopyJob* job = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] {
          for (CopyJob* j : jobs_)
            if (j->next.load() < j->chunks->size()) { job = j; return true; }
          return false;
        });
        ++job->workers;   // the owner does not return (and destroy the job) while a worker still points at it
      }
      work_on(job);
      {
        std::lock_guard<std::mutex> lk(mu_);
        --job->workers;
        done_cv_.notify_all();
      }
    }
  }
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  std::vector<CopyJob*> jobs_;
  std::vector<std::thread> threads_;
};

void parallel_copy(const std::vector<CopyPiece>& pieces) {
  constexpr size_t kChunk = size_t(4) << 20;
  std::vector<CopyPiece> chunks;
  size_t total = 0;
  for (const CopyPiece& p : pieces) {
    total += p.bytes * p.rows;
    if (p.rows <= 1 || p.src_pitch == p.bytes) {   // one contiguous run
      const size_t n = p.bytes * p.rows;
      for (size_t o = 0; o < n; o += kChunk) chunks.push_back({p.dst + o, p.src + o, std::min(kChunk, n - o)});
    } else {                                       // strided rows: chunks of whole rows
      const size_t per = std::max<size_t>(1, kChunk / p.bytes);
      for (size_t r = 0; r < p.rows; r += per)
        chunks.push_back({p.dst + r * p.bytes, p.src + r * p.src_pitch, p.bytes, std::min(per, p.rows - r), p.src_pitch});
    }
  }
  auto copy_chunk = [](const CopyPiece& c) {
    if (c.rows <= 1) { memcpy(c.dst, c.src, c.bytes); return; }
    for (size_t r = 0; r < c.rows; ++r) memcpy(c.dst + r * c.bytes, c.src + r * c.src_pitch, c.bytes);
  };
  if (total < (size_t(8) << 20) || chunks.size() <= 1 || CopyWorkers::get().size() == 0) {
    for (const CopyPiece& c : chunks) copy_chunk(c);
    return;
  }
  static const bool use_pool = [] { const char* v = getenv("CLIPB200_COPY_POOL"); return v == nullptr || atoi(v) != 0; }();
  if (!use_pool) {   // A/B switch: round 2's first version, up to 15 freshly spawned threads per call
    std::atomic<size_t> next(0);
    auto work = [&]() {
      for (size_t i; (i = next.fetch_add(1)) < chunks.size();) copy_chunk(chunks[i]);
    };
    std::vector<std::thread> pool;
    for (size_t t = 1; t < std::min<size_t>(16, chunks.size()); ++t) pool.emplace_back(work);
    work();
    for (std::thread& t : pool) t.join();
    return;
  }
  // a job = this call's chunk list; the process-wide workers and the calling thread pull chunks from it
  CopyJob job;
  job.chunks = &chunks;
  job.run = +[](const CopyPiece& c) {
    if (c.rows <= 1) { memcpy(c.dst, c.src, c.bytes); return; }
    for (size_t r = 0; r < c.rows; ++r) memcpy(c.dst + r * c.bytes, c.src + r * c.src_pitch, c.bytes);
  };
  CopyWorkers::get().run(&job);
}
}  // namespace

// Takes as many of the `count` images as fit one staging group (at least one), stages the rows and columns the resize
// reads through pinned memory (process-wide copy workers), uploads them and their coefficient tables on the copy-in stream
// and resizes them into d_dst[i * S*S*3] with two launches on the resize stream (so the next group's upload does not
// queue behind them); the caller's compute stream keeps running the previous micro-batch's tower.
Status Engine::ResizeGroupToDevice(const uint8_t* const* imgs, const int32_t* widths, const int32_t* heights, int count,
                                   const clipb200_preproc* pp, uint8_t* d_dst, int* consumed) {
  const size_t px = static_cast<size_t>(S_) * S_ * 3;
  const bool squash = pp->resize_mode == 1;
  std::vector<ResizeJob> jobs;
  std::vector<int32_t> arena;
  std::vector<CopyPiece> pieces;
  std::map<AxisKey, int> placed;  // axis -> word offset of its `start` table in this group's arena
  size_t src_bytes = 0, tmp_bytes = 0;
  int max_rows = 0;
  auto place_axis = [&](int in_size, double in0, double in1, int* start, int* size, int* w, int* window, int* precision,
                        int* first, int* last) {
    const AxisEntry& e = GetAxis(in_size, in0, in1, pp->interpolation);
    AxisKey key;
    key.in_size = in_size;
    key.interp = pp->interpolation;
    memcpy(&key.in0_bits, &in0, 8);
    memcpy(&key.in1_bits, &in1, 8);
    auto it = placed.find(key);
    if (it == placed.end()) {
      const int off = static_cast<int>(arena.size());
      arena.insert(arena.end(), e.axis.start.begin(), e.axis.start.end());
      arena.insert(arena.end(), e.axis.size.begin(), e.axis.size.end());
      const size_t wwords = (e.axis.w.size() + 1) / 2;
      const size_t at = arena.size();
      arena.resize(at + wwords, 0);
      memcpy(&arena[at], e.axis.w.data(), e.axis.w.size() * 2);
      it = placed.emplace(key, off).first;
    }
    *start = it->second;
    *size = it->second + S_;
    *w = it->second + 2 * S_;
    *window = e.axis.window;
    *precision = e.axis.precision;
    *first = e.first;
    *last = e.last;
  };
  int n = 0;
  for (; n < count; ++n) {
    const int W = widths[n], H = heights[n];
    if (imgs[n] == nullptr) return Status::Err(CLIPB200_ERR_INVALID_ARG, "null image pointer");
    if (W <= 0 || H <= 0 || W > 32768 || H > 32768) return Status::Err(CLIPB200_ERR_INVALID_ARG, "bad image size");
    const size_t bytes = static_cast<size_t>(W) * H * 3;
    if (n > 0 && (src_bytes + bytes > rs_group_bytes_ || arena.size() > (size_t(8) << 20))) break;
    ResizeJob j;
    memset(&j, 0, sizeof(j));
    j.W = W;
    j.H = H;
    j.src_off = static_cast<long long>(src_bytes);
    j.dst_off = static_cast<long long>(static_cast<size_t>(n) * px);
    j.tmp_off = static_cast<long long>(tmp_bytes);
    size_t first_row = 0, n_rows = static_cast<size_t>(H);  // source rows that have to travel
    size_t first_col = 0, n_cols = static_cast<size_t>(W);  // ... and the columns of each of them
    j.pitch = W;
    if (W == S_ && H == S_) {
      j.mode = 2;  // the convolution is the identity at the model resolution (and the crop box is the whole image)
    } else {
      double left, top, cw, ch;
      resize_crop_box(W, H, S_, squash, &left, &top, &cw, &ch);
      if (pp->interpolation > 1) {
        j.mode = 1;  // ResizeAlg::Nearest (vision.rs:179)
        j.left = left; j.top = top; j.sx = cw / S_; j.sy = ch / S_;
      } else {
        j.mode = 0;
        int xf, xl, yf, yl;
        place_axis(W, left, left + cw, &j.xstart, &j.xsize, &j.xw, &j.xwindow, &j.xprecision, &xf, &xl);
        place_axis(H, top, top + ch, &j.ystart, &j.ysize, &j.yw, &j.ywindow, &j.yprecision, &yf, &yl);
        // only the rows the vertical pass reads and the columns the horizontal pass reads cross PCIe (a centre crop
        // of a portrait photo skips rows, of a landscape photo columns: 44 % of a 16:9 frame)
        first_row = static_cast<size_t>(yf);
        n_rows = static_cast<size_t>(std::max(yl - yf, 0));
        j.y_first = yf;  // staged row r is source row yf + r
        j.rows = static_cast<int>(n_rows);
        static const bool full_rows = [] { const char* v = getenv("CLIPB200_STAGE_FULL_ROWS"); return v != nullptr && atoi(v) != 0; }();
        if (!full_rows) {   // CLIPB200_STAGE_FULL_ROWS=1: A/B switch, stage whole rows as round 2's first version did
          first_col = static_cast<size_t>(xf);
          n_cols = static_cast<size_t>(std::max(xl - xf, 0));
          j.x_first = xf;
          j.pitch = static_cast<int>(n_cols);
        }
        tmp_bytes += n_rows * S_ * 3;
        max_rows = std::max(max_rows, j.rows);
      }
    }
    const size_t staged = n_rows * n_cols * 3;
    pieces.push_back({nullptr, imgs[n] + (first_row * W + first_col) * 3, n_cols * 3, n_rows, static_cast<size_t>(W) * 3});
    src_bytes += (staged + 15) & ~size_t(15);
    jobs.push_back(j);
  }
  *consumed = n;
  ResizeStage* st = &rs_stage_[rs_groups_++ % static_cast<uint64_t>(rs_stages_)];
  if (st->in_flight) CUDA_RET(cudaEventSynchronize(st->free_ev), "wait resize staging");
  RET_IF_ERR(GrowStage(st, src_bytes, tmp_bytes, arena.size(), jobs.size()));
  {
    size_t off = 0;
    for (size_t i = 0; i < pieces.size(); ++i) {
      pieces[i].dst = st->h_src + off;
      off += (pieces[i].bytes * pieces[i].rows + 15) & ~size_t(15);
    }
    parallel_copy(pieces);
  }
  memcpy(st->h_jobs, jobs.data(), jobs.size() * sizeof(ResizeJob));
  if (!arena.empty()) memcpy(st->h_arena, arena.data(), arena.size() * 4);
  ProfBegin(PC_H2D, copy_in_);
  cudaError_t e = cudaMemcpyAsync(st->d_src, st->h_src, src_bytes, cudaMemcpyHostToDevice, copy_in_);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(st->d_jobs, st->h_jobs, jobs.size() * sizeof(ResizeJob), cudaMemcpyHostToDevice, copy_in_);
  if (e == cudaSuccess && !arena.empty())
    e = cudaMemcpyAsync(st->d_arena, st->h_arena, arena.size() * 4, cudaMemcpyHostToDevice, copy_in_);
  ProfEnd(PC_H2D, copy_in_);
  CUDA_RET(e, "H2D copy");
  // the resize kernels run on their own stream: the next group's upload does not wait for them
  CUDA_RET(cudaEventRecord(st->h2d_ev, copy_in_), "record");
  CUDA_RET(cudaStreamWaitEvent(resize_, st->h2d_ev, 0), "wait upload");
  ProfBegin(PC_PRE, resize_);
  e = launch_resize_batched(st->d_src, st->d_jobs, st->d_arena, n, S_, max_rows, st->d_tmp, d_dst, resize_);
  ProfEnd(PC_PRE, resize_);
  if (max_rows > 0) ++launch_count;  // two launches per group (ProfBegin counted one)
  CUDA_RET(e, "resize");
  CUDA_RET(cudaEventRecord(st->free_ev, resize_), "record");
  st->in_flight = true;
  return Status::OK();
}

Status Engine::ResizeRgb8(const uint8_t* img, int width, int height, const clipb200_preproc* pp, uint8_t* out) {
  if (kind != CLIPB200_KIND_VISION) return Status::Err(CLIPB200_ERR_INVALID_ARG, "not a vision engine");
  if (img == nullptr || out == nullptr || pp == nullptr) return Status::Err(CLIPB200_ERR_INVALID_ARG, "null buffer");
  CUDA_RET(cudaSetDevice(device), "cudaSetDevice");
  uint8_t* dst = static_cast<uint8_t*>(d_in_[0]);
  const int32_t w = width, h = height;
  int consumed = 0;
  RET_IF_ERR(ResizeGroupToDevice(&img, &w, &h, 1, pp, dst, &consumed));
  CUDA_RET(cudaMemcpyAsync(out, dst, static_cast<size_t>(S_) * S_ * 3, cudaMemcpyDeviceToHost, resize_), "D2H copy");
  CUDA_RET(cudaStreamSynchronize(resize_), "sync");
  return Status::OK();
}

// embed_images(&[DynamicImage]) for photos of any size (vision.rs:102-117 with the resize of :164-198 on the GPU).
// Same three-stream pipeline as RunPipelined: while the tower of micro-batch i runs on the compute stream, the host
// stages micro-batch i+1's photos group by group and the copy-in stream uploads and resizes them into the other slot.
Status Engine::VisionEmbedRgb8Var(const uint8_t* const* imgs, const int32_t* widths, const int32_t* heights, int64_t batch,
                                  const clipb200_preproc* pp, float* out) {
  if (kind != CLIPB200_KIND_VISION) return Status::Err(CLIPB200_ERR_INVALID_ARG, "not a vision engine");
  if (batch <= 0) return Status::Err(CLIPB200_ERR_INVALID_ARG, "Empty batch");
  if (imgs == nullptr || widths == nullptr || heights == nullptr || out == nullptr || pp == nullptr)
    return Status::Err(CLIPB200_ERR_INVALID_ARG, "null buffer");
  CUDA_RET(cudaSetDevice(device), "cudaSetDevice");
  RET_IF_ERR(SetPreproc(pp));
  const size_t px = static_cast<size_t>(S_) * S_ * 3;
  cudaPointerAttributes attr;
  const bool out_pinned = cudaPointerGetAttributes(&attr, out) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  // Photos are tens of MB each: staging + PCIe of a micro-batch takes as long as its tower, so the micro-batch is kept
  // small enough (<= 64 images) that even a call with a few dozen photos overlaps the two.
  static const int64_t photo_mb = [] { const char* v = getenv("CLIPB200_PHOTO_MB"); return v != nullptr && atoi(v) > 0 ? atoi(v) : 64; }();
  const int64_t mb = std::min<int64_t>(mb_, photo_mb);
  // Equal micro-batches.  (Measured and dropped: halving the tail of the call so that the last, exposed tower is a small
  // one — 112 photos as 64 + 24 + 12 + 12 — is slower, 914 -> 804 img/s on SO400M: every micro-batch costs the staging
  // host thread a tower's worth of kernel launches.  profiles/r02z_photos.md)
  std::vector<std::pair<int64_t, int>> sched;   // (first image, count)
  for (int64_t at = 0; at < batch; at += mb) sched.emplace_back(at, static_cast<int>(std::min<int64_t>(mb, batch - at)));
  const int64_t steps = static_cast<int64_t>(sched.size());
  Status st = Status::OK();
  for (int64_t s = 0; s < steps && st.ok(); ++s) {
    const int slot = static_cast<int>(s & 1);
    const int64_t at = sched[static_cast<size_t>(s)].first;
    const int n = sched[static_cast<size_t>(s)].second;
    if (s >= 2) CUDA_RET(cudaStreamWaitEvent(resize_, in_consumed_[slot], 0), "wait consumed");   // resize_ writes the slot
    uint8_t* d_slot = static_cast<uint8_t*>(d_in_[slot]);
    for (int i = 0; i < n && st.ok();) {
      const int64_t g = at + i;
      int consumed = 0;
      st = ResizeGroupToDevice(imgs + g, widths + g, heights + g, n - i, pp, d_slot + static_cast<size_t>(i) * px, &consumed);
      i += consumed;
    }
    if (!st.ok()) break;
    CUDA_RET(cudaEventRecord(in_ready_[slot], resize_), "record");
    BindLane(slot);
    CUDA_RET(cudaStreamWaitEvent(compute_, in_ready_[slot], 0), "wait input");
    if (s >= 2) CUDA_RET(cudaStreamWaitEvent(compute_, out_copied_[slot], 0), "wait output slot");
    st = ForwardSlot(0, n, slot);
    if (!st.ok()) break;
    CUDA_RET(cudaEventRecord(in_consumed_[slot], compute_), "record");
    CUDA_RET(cudaEventRecord(out_ready_[slot], compute_), "record");
    CUDA_RET(cudaStreamWaitEvent(copy_out_, out_ready_[slot], 0), "wait output");
    ProfBegin(PC_D2H, copy_out_);
    cudaError_t e = cudaMemcpyAsync(out_pinned ? out + static_cast<size_t>(at) * E_ : h_out_[slot], d_out_[slot],
                                    static_cast<size_t>(n) * E_ * 4, cudaMemcpyDeviceToHost, copy_out_);
    ProfEnd(PC_D2H, copy_out_);
    CUDA_RET(e, "D2H copy");
    CUDA_RET(cudaEventRecord(out_copied_[slot], copy_out_), "record");
    if (!out_pinned && s >= 1) {  // drain the previous step's output while this one runs
      const int ps = static_cast<int>((s - 1) & 1);
      CUDA_RET(cudaEventSynchronize(out_copied_[ps]), "wait D2H");
      memcpy(out + static_cast<size_t>(sched[static_cast<size_t>(s - 1)].first) * E_, h_out_[ps],
             static_cast<size_t>(sched[static_cast<size_t>(s - 1)].second) * E_ * 4);
    }
  }
  BindLane(0);
  Status sync = Synchronize();
  if (!st.ok()) return st;
  RET_IF_ERR(sync);
  if (!out_pinned) {
    const int64_t s = steps - 1;
    memcpy(out + static_cast<size_t>(sched[static_cast<size_t>(s)].first) * E_, h_out_[s & 1],
           static_cast<size_t>(sched[static_cast<size_t>(s)].second) * E_ * 4);
  }
  return Status::OK();
}

Status Engine::VisionEmbedF32(const float* nchw, int64_t batch, float* out) {
  if (kind != CLIPB200_KIND_VISION) return Status::Err(CLIPB200_ERR_INVALID_ARG, "not a vision engine");
  if (batch <= 0) return Status::Err(CLIPB200_ERR_INVALID_ARG, "Empty batch");
  if (nchw == nullptr || out == nullptr) return Status::Err(CLIPB200_ERR_INVALID_ARG, "null buffer");
  CUDA_RET(cudaSetDevice(device), "cudaSetDevice");
  const size_t slot = static_cast<size_t>(mb_) * 3 * S_ * S_ * 4;
  for (int i = 0; i < 2; ++i) {
    if (d_in_f32_[i] == nullptr) CUDA_RET(cudaMalloc(reinterpret_cast<void**>(&d_in_f32_[i]), slot), "f32 input slot");
    if (h_in_f32_[i] == nullptr)
      CUDA_RET(cudaHostAlloc(reinterpret_cast<void**>(&h_in_f32_[i]), slot, cudaHostAllocDefault), "f32 pinned slot");
  }
  return RunPipelined<float>(nchw, batch, static_cast<size_t>(3) * S_ * S_, out, false, 1);
}

Status Engine::TextEmbed(const int64_t* ids, int64_t batch, int64_t ctx, float* out, bool device_buffers) {
  if (kind != CLIPB200_KIND_TEXT) return Status::Err(CLIPB200_ERR_INVALID_ARG, "not a text engine");
  if (batch <= 0) return Status::Err(CLIPB200_ERR_INVALID_ARG, "Empty batch");
  if (ids == nullptr || out == nullptr) return Status::Err(CLIPB200_ERR_INVALID_ARG, "null buffer");
  if (ctx != T_)
    return Status::Err(CLIPB200_ERR_INVALID_ARG, "input_ids has context length " + std::to_string(ctx) +
                                                     ", the graph expects " + std::to_string(T_));
  return RunPipelined<int64_t>(ids, batch, static_cast<size_t>(T_), out, device_buffers, 2);
}

}  // namespace clipb200
