// FastViT / MobileCLIP2 hybrid trunk on the engine: loader (timm names after `reparameterize_model`, eval BatchNorm
// folded; reference pull_onnx.py:110-116) and forward pass.  Activations are NHWC, so every 1x1 convolution is a
// plain [pixels, Cin] x [Cout, Cin]^T GEMM on the tcgen05 kernel (bias / GELU / layer-scale + residual fused in its
// epilogue); depthwise convolutions, SE gates and pooling are the kernels in conv_kernels.cu; the attention blocks
// of the last stage reuse the attention kernels (64 tokens, head_dim 32).
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "attn_sm100.cuh"
#include "conv_kernels.cuh"
#include "engine.h"
#include "fused_mlp_sm100.cuh"
#include "gemm_sm100.cuh"
#include "kernels.cuh"

namespace clipb200 {

#define RET_IF_ERR(expr)      \
  do {                        \
    Status s_ = (expr);       \
    if (!s_.ok()) return s_;  \
  } while (0)
#define CUDA_RET(expr, what)           \
  do {                                 \
    Status s_ = Check((expr), what);   \
    if (!s_.ok()) return s_;           \
  } while (0)

static std::vector<int> parse_ints(const std::string& s) {
  std::vector<int> v;
  size_t pos = 0;
  while (pos < s.size()) {
    size_t e = s.find(',', pos);
    if (e == std::string::npos) e = s.size();
    v.push_back(atoi(s.substr(pos, e - pos).c_str()));
    pos = e + 1;
  }
  return v;
}

// conv weight [Cout, cin_g, k, k] -> [cin_g*k*k][Cout] fp32 (tap-major, channel contiguous), bias [Cout]
Status Engine::UploadConv(const OnnxModel& m, const std::string& name, int cout, int cin_g, int k, ConvW* out) {
  std::vector<float> w, b;
  RET_IF_ERR(HostF32(m, name + ".weight", static_cast<int64_t>(cout) * cin_g * k * k, &w));
  RET_IF_ERR(HostF32(m, name + ".bias", cout, &b));
  const int taps = cin_g * k * k;
  std::vector<float> r(static_cast<size_t>(taps) * cout);
  for (int oc = 0; oc < cout; ++oc)
    for (int t = 0; t < taps; ++t) r[static_cast<size_t>(t) * cout + oc] = w[static_cast<size_t>(oc) * taps + t];
  RET_IF_ERR(UploadHostF32(r.data(), r.size(), &out->w));
  RET_IF_ERR(UploadHostF32(b.data(), b.size(), &out->b));
  out->cout = cout;
  out->k = k;
  return Status::OK();
}

Status Engine::UploadSe(const OnnxModel& m, const std::string& name, int C, SeW* out) {
  const OnnxTensor* t = m.find(name + ".fc1.weight");
  if (t == nullptr) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "initializer '" + name + ".fc1.weight' not found in graph");
  const int R = static_cast<int>(t->dims[0]);
  RET_IF_ERR(UploadF32(m, name + ".fc1.weight", static_cast<int64_t>(R) * C, &out->w1));
  RET_IF_ERR(UploadF32(m, name + ".fc1.bias", R, &out->b1));
  RET_IF_ERR(UploadF32(m, name + ".fc2.weight", static_cast<int64_t>(C) * R, &out->w2));
  RET_IF_ERR(UploadF32(m, name + ".fc2.bias", C, &out->b2));
  out->C = C;
  out->R = R;
  return Status::OK();
}

Status Engine::LoadFastVit(const OnnxModel& m) {
  fastvit_ = true;
  if (const char* env = getenv("CLIPB200_FUSED_MLP")) fused_mlp_ = atoi(env) != 0;
  CUDA_RET(fused_mlp_configure_device(), "configure fused ConvMlp kernels");
  family_ = "fastvit";
  const std::string pre = "model.visual.trunk";
  const OnnxTensor* s0 = m.find(pre + ".stem.0.reparam_conv.weight");
  if (s0->dims.size() != 4 || s0->dims[1] != 3 || s0->dims[2] != 3)
    return Status::Err(CLIPB200_ERR_UNSUPPORTED, "FastViT stem must be a 3x3 conv on 3 channels");
  const int d0 = static_cast<int>(s0->dims[0]);
  // stage widths / depths: metadata if present, else probe the initializer names
  std::vector<int> dims = parse_ints(m.meta("clipb200.dims")), depths = parse_ints(m.meta("clipb200.depths"));
  if (dims.empty()) {
    for (int i = 0; i < 8; ++i) {
      const std::string st = pre + ".stages." + std::to_string(i);
      const OnnxTensor* t = m.find(st + ".blocks.0.mlp.fc2.weight");
      if (t == nullptr) break;
      dims.push_back(static_cast<int>(t->dims[0]));
      int d = 0;
      while (m.has(st + ".blocks." + std::to_string(d) + ".mlp.fc2.weight")) ++d;
      depths.push_back(d);
    }
  }
  if (dims.empty() || dims.size() != depths.size() || dims[0] != d0)
    return Status::Err(CLIPB200_ERR_UNSUPPORTED, "cannot determine FastViT stage layout");
  // image size: metadata, else the declared input shape [batch, 3, S, S] of a real export, else MobileCLIP2's 256
  const std::string isz = m.meta("clipb200.image_size");
  S_ = isz.empty() ? 256 : atoi(isz.c_str());
  if (isz.empty() && m.input_infos.size() == 1 && m.input_infos[0].dims.size() == 4 && m.input_infos[0].dims[2] > 0 &&
      m.input_infos[0].dims[2] == m.input_infos[0].dims[3])
    S_ = static_cast<int>(m.input_infos[0].dims[2]);
  image_size = S_;
  if (S_ % 32 != 0) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "FastViT image size must be a multiple of 32");
  act_ = ACT_GELU_ERF;
  RET_IF_ERR(UploadConv(m, pre + ".stem.0.reparam_conv", d0, 3, 3, &fv_stem0_));
  RET_IF_ERR(UploadConv(m, pre + ".stem.1.reparam_conv", d0, 1, 3, &fv_stem1_));
  RET_IF_ERR(UploadLinear(m, pre + ".stem.2.reparam_conv.weight", pre + ".stem.2.reparam_conv.bias", d0, d0, false, &fv_stem2_));
  int prev = d0;
  fv_stages_.resize(dims.size());
  for (size_t i = 0; i < dims.size(); ++i) {
    FvStage& st = fv_stages_[i];
    const int C = dims[i];
    st.C = C;
    const std::string sp = pre + ".stages." + std::to_string(i);
    if (i > 0) {
      if (C % prev != 0 || C / prev != 2) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "FastViT stages must double the width");
      st.down = true;
      RET_IF_ERR(UploadConv(m, sp + ".downsample.proj.0.reparam_conv", C, 1, 7, &st.down_dw));
      st.down_se = m.has(sp + ".downsample.proj.0.se.fc1.weight");
      if (st.down_se) RET_IF_ERR(UploadSe(m, sp + ".downsample.proj.0.se", C, &st.se));
      RET_IF_ERR(UploadLinear(m, sp + ".downsample.proj.1.reparam_conv.weight", sp + ".downsample.proj.1.reparam_conv.bias", C, C,
                              false, &st.down_pw));
    }
    st.cpe = m.has(sp + ".pos_emb.reparam_conv.weight");
    if (st.cpe) RET_IF_ERR(UploadConv(m, sp + ".pos_emb.reparam_conv", C, 1, 7, &st.cpe_dw));
    st.blocks.resize(depths[i]);
    for (int j = 0; j < depths[i]; ++j) {
      FvBlock& b = st.blocks[j];
      const std::string bp = sp + ".blocks." + std::to_string(j);
      const OnnxTensor* fc1 = m.find(bp + ".mlp.fc1.weight");
      if (fc1 == nullptr) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "initializer '" + bp + ".mlp.fc1.weight' not found in graph");
      const int hidden = static_cast<int>(fc1->dims[0]);
      RET_IF_ERR(UploadConv(m, bp + ".mlp.conv.conv", C, 1, 7, &b.mlp_dw));
      RET_IF_ERR(UploadLinear(m, bp + ".mlp.fc1.weight", bp + ".mlp.fc1.bias", hidden, C, false, &b.fc1));
      RET_IF_ERR(UploadLinear(m, bp + ".mlp.fc2.weight", bp + ".mlp.fc2.bias", C, hidden, false, &b.fc2));
      b.attn = m.has(bp + ".token_mixer.qkv.weight");
      if (!b.attn) {
        RET_IF_ERR(UploadConv(m, bp + ".token_mixer.reparam_conv", C, 1, 3, &b.mixer));
        RET_IF_ERR(UploadF32(m, bp + ".layer_scale.gamma", C, &b.gamma));
      } else {
        if (C % 32 != 0) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "attention stage width must be a multiple of 32");
        // eval BatchNorm in front of the attention folds into qkv: W' = W diag(s), b' = W t
        std::vector<float> nw, nb, mean, var, wq;
        RET_IF_ERR(HostF32(m, bp + ".norm.weight", C, &nw));
        RET_IF_ERR(HostF32(m, bp + ".norm.bias", C, &nb));
        RET_IF_ERR(HostF32(m, bp + ".norm.running_mean", C, &mean));
        RET_IF_ERR(HostF32(m, bp + ".norm.running_var", C, &var));
        RET_IF_ERR(HostF32(m, bp + ".token_mixer.qkv.weight", static_cast<int64_t>(3) * C * C, &wq));
        std::vector<float> sc(C), sh(C), bq(3 * C);
        for (int c = 0; c < C; ++c) {
          sc[c] = nw[c] / sqrtf(var[c] + 1e-5f);
          sh[c] = nb[c] - mean[c] * sc[c];
        }
        for (int o = 0; o < 3 * C; ++o) {
          double acc = 0.0;
          float* row = &wq[static_cast<size_t>(o) * C];
          for (int c = 0; c < C; ++c) {
            acc += static_cast<double>(row[c]) * sh[c];
            row[c] *= sc[c];
          }
          bq[o] = static_cast<float>(acc);
        }
        RET_IF_ERR(UploadLinearFromHost(wq.data(), 3 * C, C, bq.data(), &b.qkv));
        RET_IF_ERR(UploadLinear(m, bp + ".token_mixer.proj.weight", bp + ".token_mixer.proj.bias", C, C, false, &b.proj));
        RET_IF_ERR(UploadF32(m, bp + ".layer_scale_1.gamma", C, &b.gamma1));
        RET_IF_ERR(UploadF32(m, bp + ".layer_scale_2.gamma", C, &b.gamma2));
      }
    }
    prev = C;
  }
  const int cf = 2 * prev;
  RET_IF_ERR(UploadConv(m, pre + ".final_conv.reparam_conv", cf, 1, 3, &fv_final_));
  RET_IF_ERR(UploadSe(m, pre + ".final_conv.se", cf, &fv_final_se_));
  const OnnxTensor* hw = m.find(pre + ".head.fc.weight");
  if (hw == nullptr) return Status::Err(CLIPB200_ERR_UNSUPPORTED, "FastViT head.fc.weight not found");
  E_ = static_cast<int>(hw->dims[0]);
  embed_dim = E_;
  RET_IF_ERR(UploadLinear(m, pre + ".head.fc.weight", pre + ".head.fc.bias", E_, cf, false, &head_));
  D_ = prev;
  mlp_ = 3 * prev;
  H_ = prev / 32;
  hd_ = 32;
  T_ = 1;
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&lut_), 3 * 256 * 4));
  return Status::OK();
}

Status Engine::AllocFastVitWorkspace() {
  const size_t P0 = static_cast<size_t>(S_ / 4) * (S_ / 4);
  size_t max_pc = 0, max_hidden = 0, max_c = 0, max_qkv = 16;
  size_t P = P0;
  for (size_t i = 0; i < fv_stages_.size(); ++i) {
    if (i > 0) P /= 4;
    const size_t C = fv_stages_[i].C;
    max_pc = std::max(max_pc, P * C);
    for (const FvBlock& b : fv_stages_[i].blocks) {
      max_hidden = std::max(max_hidden, P * static_cast<size_t>(b.fc1.N));
      if (b.attn) max_qkv = std::max(max_qkv, P * 3 * C);  // 5-stage trunks (MCi3 / MCi4) have two attention stages
    }
    max_c = std::max(max_c, C);
  }
  const size_t cf = 2 * static_cast<size_t>(fv_stages_.back().C);
  max_pc = std::max(max_pc, P * cf);
  max_c = std::max(max_c, cf);
  const size_t mb = static_cast<size_t>(mb_);
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&fv_xa_), mb * max_pc * 4));
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&fv_xb_), mb * max_pc * 4));
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&fv_tmp_), mb * max_pc * 4));
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&fv_s_), mb * max_c * 4));
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&fv_gate_), mb * max_c * 4));
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&fv_stem_out_), mb * static_cast<size_t>(S_ / 2) * (S_ / 2) * fv_stem0_.cout * 2));
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&h_), mb * max_pc * 2));
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&mlpbuf_), mb * max_hidden * 2));
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&qkv_), mb * max_qkv * 2));
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&pooled_), mb * cf * 2));
  RET_IF_ERR(DevAlloc(reinterpret_cast<void**>(&proj_out_), mb * static_cast<size_t>(E_) * 4));
  return Status::OK();
}

// depthwise conv with its own profile class: device time + algorithmic bytes (input once + output once), so the
// bench can state the achieved GB/s of the conv stages next to the GEMM class's TFLOP/s
cudaError_t Engine::DwConv(const void* in, bool in_bf16, int n, int H, int W, int Cin, int K, int stride, int mult,
                           const float* w, const float* bias, bool gelu, void* out, bool out_bf16) {
  ProfBegin(PC_CONV, compute_);
  cudaError_t e = launch_dwconv(in, in_bf16, n, H, W, Cin, K, stride, mult, w, bias, gelu, out, out_bf16, compute_);
  ProfEnd(PC_CONV, compute_);
  if (profile_) {
    const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
    prof_acc_.conv_bytes += static_cast<double>(n) * H * W * Cin * (in_bf16 ? 2 : 4) +
                            static_cast<double>(n) * Ho * Wo * Cin * mult * (out_bf16 ? 2 : 4);
  }
  return e;
}

// ConvMlp: x += gamma * fc2(gelu(fc1(dwconv7x7(x))))
Status Engine::FvMlp(const FvBlock& b, float* cur, int n, int Hh, int C, const float* gamma) {
  const int rows = n * Hh * Hh;
  cudaError_t e = DwConv(cur, false, n, Hh, Hh, C, 7, 1, 1, b.mlp_dw.w, b.mlp_dw.b, false, h_, true);
  CUDA_RET(e, "depthwise 7x7");
  if (fused_mlp_ && fused_mlp_supported(C, b.fc1.N) && b.fc1.ldk == C && b.fc2.ldk == b.fc1.N) {
    // fc1 -> GELU -> fc2 -> layer scale -> residual in one kernel: the 3C-wide hidden activation stays on the SM
    ProfBegin(PC_GEMM, compute_);
    e = fused_mlp(h_, C, b.fc1.w, b.fc1.ldk, b.fc1.b, b.fc2.w, b.fc2.ldk, b.fc2.b, gamma, cur, C, rows, C, b.fc1.N, num_sms_, compute_);
    ProfEnd(PC_GEMM, compute_);
    if (profile_) prof_acc_.gemm_flops += 4.0 * rows * static_cast<double>(C) * b.fc1.N;
    return Check(e, "fused ConvMlp");
  }
  GemmEpilogue e1;
  e1.out_bf16 = mlpbuf_;
  e1.ldc = b.fc1.N;
  e1.act = ACT_GELU_ERF;
  RET_IF_ERR(Gemm(h_, C, b.fc1, rows, EPI_BF16, &e1));
  GemmEpilogue e2;
  e2.out_f32 = cur;
  e2.ldc = C;
  e2.gamma = gamma;
  RET_IF_ERR(Gemm(mlpbuf_, b.fc1.N, b.fc2, rows, EPI_RESID, &e2));
  return Status::OK();
}

Status Engine::ForwardFastVit(int n, const uint8_t* d_u8, const float* d_f32, float* d_out) {
  cudaError_t e;
  const int d0 = fv_stem0_.cout;
  int Hh = S_ / 2;
  ProfBegin(PC_PRE, compute_);
  e = launch_stem_conv3x3_s2(d_u8, d_f32, lut_, n, S_, d0, fv_stem0_.w, fv_stem0_.b, fv_stem_out_, compute_);
  ProfEnd(PC_PRE, compute_);
  CUDA_RET(e, "stem conv");
  e = DwConv(fv_stem_out_, true, n, Hh, Hh, d0, 3, 2, 1, fv_stem1_.w, fv_stem1_.b, true, h_, true);
  CUDA_RET(e, "stem depthwise");
  Hh /= 2;
  float* cur = fv_xa_;
  float* other = fv_xb_;
  {
    GemmEpilogue ep;
    ep.out_f32 = cur;
    ep.ldc = d0;
    ep.act = ACT_GELU_ERF;
    RET_IF_ERR(Gemm(h_, d0, fv_stem2_, n * Hh * Hh, EPI_F32, &ep));
  }
  int C = d0;
  for (size_t i = 0; i < fv_stages_.size(); ++i) {
    const FvStage& st = fv_stages_[i];
    if (st.down) {
      const int Ho = Hh / 2;
      if (st.down_se) {
              e = DwConv(cur, false, n, Hh, Hh, C, 7, 2, 2, st.down_dw.w, st.down_dw.b, false, fv_tmp_, false);
        CUDA_RET(e, "downsample depthwise");
        ProfBegin(PC_MISC, compute_);
        e = launch_gap(fv_tmp_, n, Ho * Ho, st.C, fv_s_, compute_);
        if (e == cudaSuccess) e = launch_se_mlp(fv_s_, n, st.C, st.se.R, st.se.w1, st.se.b1, st.se.w2, st.se.b2, fv_gate_, compute_);
        if (e == cudaSuccess) e = launch_scale_act(fv_tmp_, fv_gate_, n, Ho * Ho, st.C, true, h_, true, compute_);
        ProfEnd(PC_MISC, compute_);
        CUDA_RET(e, "downsample squeeze-excite");
      } else {
              e = DwConv(cur, false, n, Hh, Hh, C, 7, 2, 2, st.down_dw.w, st.down_dw.b, true, h_, true);
        CUDA_RET(e, "downsample depthwise");
      }
      Hh = Ho;
      C = st.C;
      GemmEpilogue ep;
      ep.out_f32 = cur;
      ep.ldc = C;
      ep.act = ACT_GELU_ERF;
      RET_IF_ERR(Gemm(h_, C, st.down_pw, n * Hh * Hh, EPI_F32, &ep));
    }
    if (st.cpe) {
          e = DwConv(cur, false, n, Hh, Hh, C, 7, 1, 1, st.cpe_dw.w, st.cpe_dw.b, false, other, false);
      CUDA_RET(e, "positional encoding");
      std::swap(cur, other);
    }
    const int rows = n * Hh * Hh;
    for (const FvBlock& b : st.blocks) {
      if (!b.attn) {
              e = DwConv(cur, false, n, Hh, Hh, C, 3, 1, 1, b.mixer.w, b.mixer.b, false, other, false);
        CUDA_RET(e, "token mixer");
        std::swap(cur, other);
        RET_IF_ERR(FvMlp(b, cur, n, Hh, C, b.gamma));
      } else {
        ProfBegin(PC_MISC, compute_);
        e = launch_scale_act(cur, nullptr, n, Hh * Hh, C, false, h_, true, compute_);
        ProfEnd(PC_MISC, compute_);
        CUDA_RET(e, "cast");
        GemmEpilogue eq;
        eq.out_bf16 = qkv_;
        eq.ldc = 3 * C;
        RET_IF_ERR(Gemm(h_, C, b.qkv, rows, EPI_BF16, &eq));
        ProfBegin(PC_ATTN, compute_);
        e = launch_flash_attention(qkv_, h_, n, Hh * Hh, C / 32, 32, false, compute_);
        ProfEnd(PC_ATTN, compute_);
        CUDA_RET(e, "attention");
        GemmEpilogue eo;
        eo.out_f32 = cur;
        eo.ldc = C;
        eo.gamma = b.gamma1;
        RET_IF_ERR(Gemm(h_, C, b.proj, rows, EPI_RESID, &eo));
        RET_IF_ERR(FvMlp(b, cur, n, Hh, C, b.gamma2));
      }
    }
  }
  // final_conv (dw3x3, x2 channels) -> SE -> GELU -> global average pool -> head -> L2 normalise
  const int cf = 2 * C, P = Hh * Hh;
  e = DwConv(cur, false, n, Hh, Hh, C, 3, 1, 2, fv_final_.w, fv_final_.b, false, fv_tmp_, false);
  CUDA_RET(e, "final conv");
  ProfBegin(PC_MISC, compute_);
  e = launch_gap(fv_tmp_, n, P, cf, fv_s_, compute_);
  if (e == cudaSuccess) e = launch_se_mlp(fv_s_, n, cf, fv_final_se_.R, fv_final_se_.w1, fv_final_se_.b1, fv_final_se_.w2, fv_final_se_.b2, fv_gate_, compute_);
  if (e == cudaSuccess) e = launch_scale_act(fv_tmp_, fv_gate_, n, P, cf, true, other, false, compute_);
  if (e == cudaSuccess) e = launch_gap(other, n, P, cf, fv_s_, compute_);
  if (e == cudaSuccess) e = launch_scale_act(fv_s_, nullptr, n, 1, cf, false, pooled_, true, compute_);
  ProfEnd(PC_MISC, compute_);
  CUDA_RET(e, "final squeeze-excite / pool");
  GemmEpilogue ep;
  ep.out_f32 = proj_out_;
  ep.ldc = E_;
  RET_IF_ERR(Gemm(pooled_, cf, head_, n, EPI_F32, &ep));
  ProfBegin(PC_MISC, compute_);
  e = launch_l2_normalize(proj_out_, n, E_, d_out, compute_);
  ProfEnd(PC_MISC, compute_);
  CUDA_RET(e, "l2 normalize");
  return Status::OK();
}

}  // namespace clipb200
