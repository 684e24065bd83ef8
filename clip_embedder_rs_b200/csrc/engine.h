// Engine: one (model file, GPU) pair.  Owns HBM-resident weights (bf16 GEMM operands, fp32 norms / biases /
// embeddings), an activation workspace sized for one micro-batch, two input and two output staging slots (device +
// pinned host) and three streams (H2D, compute, D2H) so that the copies of micro-batch i+1 overlap the kernels of
// micro-batch i.  It plays the role of `ort::Session` behind `OnnxSession` (reference src/onnx.rs:8-29).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/clipb200.h"
#include <map>
#include <tuple>
#include <utility>

#include "onnx_graph.h"
#include "onnx_loader.h"
#include "resize.cuh"

namespace clipb200 {

struct Status {
  int code = CLIPB200_OK;
  std::string msg;
  bool ok() const { return code == CLIPB200_OK; }
  static Status OK() { return Status(); }
  static Status Err(int c, const std::string& m) {
    Status s;
    s.code = c;
    s.msg = m;
    return s;
  }
};

struct LinearW {
  __nv_bfloat16* w = nullptr;  // [N, ldk] bf16, K contiguous
  float* b = nullptr;          // [N] or null
  int N = 0, K = 0, ldk = 0;
};
struct NormW {
  float* g = nullptr;
  float* b = nullptr;
};
struct BlockW {
  NormW ln1, ln2;
  LinearW qkv, proj, fc1, fc2;
};

enum ProfClass { PC_GEMM = 0, PC_ATTN = 1, PC_LN = 2, PC_PRE = 3, PC_MISC = 4, PC_H2D = 5, PC_D2H = 6, PC_CONV = 7 };

class Engine {
 public:
  static Status Create(const std::string& onnx_path, int device, const clipb200_opts* opts, Engine** out);
  ~Engine();

  Status VisionEmbedRgb8(const uint8_t* hwc, int64_t batch, int width, int height, const clipb200_preproc* pp,
                         float* out, bool device_buffers);
  Status VisionEmbedF32(const float* nchw, int64_t batch, float* out);
  Status PreprocessRgb8(const uint8_t* hwc, int64_t batch, int width, int height, const clipb200_preproc* pp,
                        float* out_nchw);
  Status TextEmbed(const int64_t* ids, int64_t batch, int64_t ctx, float* out, bool device_buffers);
  // arbitrary-size RGB8 images: GPU resize (vision.rs:164-198) -> normalise -> tower
  Status VisionEmbedRgb8Var(const uint8_t* const* imgs, const int32_t* widths, const int32_t* heights, int64_t batch,
                            const clipb200_preproc* pp, float* out);
  Status ResizeRgb8(const uint8_t* img, int width, int height, const clipb200_preproc* pp, uint8_t* out);

  Status RecordEvent(int slot);
  Status ElapsedMs(int a, int b, double* ms);
  Status Synchronize();
  Status ReadProfile(clipb200_profile* out, bool reset);
  Status FlushL2();

  int kind = CLIPB200_KIND_VISION;
  int device = 0;
  std::vector<std::string> input_names;
  int embed_dim = 0, image_size = 0, context_length = 0;
  int64_t weight_bytes = 0;
  int64_t launch_count = 0;

 private:
  Engine() = default;
  Status Init(const std::string& onnx_path, int device, const clipb200_opts* opts);
  Status LoadVision(const OnnxModel& m);
  Status LoadText(const OnnxModel& m);
  Status LoadBlock(const OnnxModel& m, const std::string& prefix, bool timm, BlockW* b);
  Status AllocWorkspace();
  Status UploadF32(const OnnxModel& m, const std::string& name, int64_t expect_numel, float** out);
  Status UploadLinear(const OnnxModel& m, const std::string& wname, const std::string& bname, int N, int K,
                      bool transpose, LinearW* out);
  Status HostF32(const OnnxModel& m, const std::string& name, int64_t expect_numel, std::vector<float>* out);
  Status UploadLinearFromHost(const float* w, int N, int K, const float* bias_or_null, LinearW* out);
  Status UploadHostF32(const float* src, size_t n, float** out);

  // ---- FastViT / MobileCLIP2 hybrid trunk (engine_fastvit.cu) ----
  struct ConvW {  // depthwise / stem conv, weight rearranged to [taps][Cout]
    float* w = nullptr;
    float* b = nullptr;
    int cout = 0, k = 0;
  };
  struct SeW {
    float *w1 = nullptr, *b1 = nullptr, *w2 = nullptr, *b2 = nullptr;
    int C = 0, R = 0;
  };
  struct FvBlock {
    bool attn = false;
    ConvW mixer, mlp_dw;
    LinearW fc1, fc2, qkv, proj;
    float *gamma = nullptr, *gamma1 = nullptr, *gamma2 = nullptr;
  };
  struct FvStage {
    int C = 0;
    bool down = false, down_se = false, cpe = false;
    ConvW down_dw, cpe_dw;
    SeW se;
    LinearW down_pw;
    std::vector<FvBlock> blocks;
  };
  Status LoadFastVit(const OnnxModel& m);
  Status AllocFastVitWorkspace();
  Status ForwardFastVit(int n, const uint8_t* d_u8, const float* d_f32, float* d_out);
  Status UploadConv(const OnnxModel& m, const std::string& name, int cout, int cin_g, int k, ConvW* out);
  Status UploadSe(const OnnxModel& m, const std::string& name, int C, SeW* out);
  Status FvMlp(const FvBlock& b, float* cur, int n, int Hh, int C, const float* gamma);
  cudaError_t DwConv(const void* in, bool in_bf16, int n, int H, int W, int Cin, int K, int stride, int mult,
                     const float* w, const float* bias, bool gelu, void* out, bool out_bf16);
  // CUDA graphs for small micro-batches (launch-bound regime): one instantiated graph per (mode, n, staging slot)
  Status ForwardSlot(int mode, int n, int slot);
  struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;
  };
  std::map<std::tuple<int, int, int>, GraphEntry> graphs_;
  int graph_max_n_ = 32;
  // ---- arbitrary-size images (vision.rs:164-198 on the GPU) ----
  // Coefficient tables are cached on the HOST per axis (source extent, crop interval, filter): the x and y tables of an
  // image are independent, a photo corpus has far fewer distinct axes than distinct sizes, and nothing per size lives
  // in HBM.  The cache is an LRU bounded by entries and bytes.
  struct AxisKey {
    int in_size, interp;
    uint64_t in0_bits, in1_bits;
    bool operator<(const AxisKey& o) const {
      return std::tie(in_size, interp, in0_bits, in1_bits) < std::tie(o.in_size, o.interp, o.in0_bits, o.in1_bits);
    }
  };
  struct AxisEntry {
    ResizeAxis axis;
    uint64_t tick = 0;
    size_t bytes = 0;
    int first = 0, last = 0;  // union of the source windows [first, last)
  };
  std::map<AxisKey, AxisEntry> axis_cache_;
  uint64_t axis_tick_ = 0;
  size_t axis_cache_bytes_ = 0;
  const AxisEntry& GetAxis(int in_size, double in0, double in1, int interpolation);
  // Images are resized in groups of bounded source bytes; two staging sets (pinned host + device: sources, coefficient
  // arena, jobs, intermediate) alternate so the host fills one while the copy stream drains the other.
  struct ResizeStage {
    uint8_t *h_src = nullptr, *d_src = nullptr, *d_tmp = nullptr;
    int32_t *h_arena = nullptr, *d_arena = nullptr;
    ResizeJob *h_jobs = nullptr, *d_jobs = nullptr;
    size_t src_cap = 0, tmp_cap = 0, arena_cap = 0, jobs_cap = 0;
    cudaEvent_t free_ev = nullptr, h2d_ev = nullptr;
    bool in_flight = false;
  };
  static constexpr int kResizeStages = 4;   // pinned staging groups in flight (CLIPB200_PHOTO_STAGES=2..4; measured: no effect)
  ResizeStage rs_stage_[kResizeStages];
  int rs_stages_ = 2;
  uint64_t rs_groups_ = 0;
  size_t rs_group_bytes_ = size_t(192) << 20;
  Status ResizeGroupToDevice(const uint8_t* const* imgs, const int32_t* widths, const int32_t* heights, int count,
                             const clipb200_preproc* pp, uint8_t* d_dst, int* consumed);
  Status GrowStage(ResizeStage* st, size_t src, size_t tmp, size_t arena_words, size_t jobs);
  bool fastvit_ = false;
  bool fused_mlp_ = true;   // fc1 -> GELU -> fc2 of the FastViT ConvMlp in one kernel (CLIPB200_FUSED_MLP=0: two GEMM launches)
  ConvW fv_stem0_, fv_stem1_, fv_final_;
  LinearW fv_stem2_;
  SeW fv_final_se_;
  std::vector<FvStage> fv_stages_;
  float *fv_xa_ = nullptr, *fv_xb_ = nullptr, *fv_tmp_ = nullptr, *fv_s_ = nullptr, *fv_gate_ = nullptr;
  __nv_bfloat16* fv_stem_out_ = nullptr;
  Status DevAlloc(void** p, size_t bytes);
  Status SetPreproc(const clipb200_preproc* pp);

  // forward passes on the compute stream, inputs already on the device
  Status ForwardVision(int n, const uint8_t* d_u8, const float* d_f32, float* d_out);
  Status ForwardText(int n, const int64_t* d_ids, float* d_out);
  Status Blocks(int rows, int n_seq, int T, bool causal);
  Status Gemm(const __nv_bfloat16* A, long long lda, const LinearW& w, int M, int epi, struct GemmEpilogue* ep);
  Status Check(cudaError_t e, const char* what);

  void ProfBegin(int cls, cudaStream_t st);
  void ProfEnd(int cls, cudaStream_t st);

  // architecture
  std::string family_;
  std::string graph_note_;  // why the graph recogniser declined (reported if name binding fails too)
  int S_ = 0, P_ = 0, G_ = 0, Tp_ = 0, T_ = 0, D_ = 0, L_ = 0, H_ = 0, hd_ = 0, mlp_ = 0, act_ = 0, E_ = 0;
  int K_ = 0, Kp_ = 0, vocab_ = 0;
  float eps_ = 1e-5f;
  bool has_cls_ = false, causal_ = false, pool_map_ = false, pool_argmax_ = false;
  int num_sms_ = 148;
  int mb_ = 0;  // micro-batch (sequences)
  bool profile_ = false;

  // weights
  LinearW patch_;  // [D, Kp]
  float* pos_ = nullptr;      // vision [T, D] / text [ctx, D]
  float* cls_row_ = nullptr;  // [D] = class_embedding + pos[0]
  NormW ln_pre_, ln_post_;    // ln_post_: vision ln_post / trunk.norm, text ln_final
  std::vector<BlockW> blocks_;
  LinearW head_;              // visual.proj^T / text_projection
  // MAP pool
  float* map_q_ = nullptr;    // [D] scaled query
  LinearW map_kv_, map_proj_, map_fc1_, map_fc2_;
  NormW map_norm_;
  float* tok_emb_ = nullptr;  // [vocab, D] fp32
  float* lut_ = nullptr;      // [3][256] normalisation LUT
  float lut_mean_[3] = {0, 0, 0}, lut_std_[3] = {0, 0, 0};
  bool lut_valid_ = false;

  // workspace (device)
  __nv_bfloat16 *patches_ = nullptr, *h_ = nullptr, *qkv_ = nullptr, *mlpbuf_ = nullptr, *pooled_ = nullptr,
                *yh_ = nullptr, *ymlp_ = nullptr, *vt_ = nullptr;
  bool attn_vt_ = false;
  float *x_ = nullptr, *y_ = nullptr, *proj_out_ = nullptr;
  int* row_map_ = nullptr;
  int* err_flag_ = nullptr;
  void* d_in_[2] = {nullptr, nullptr};
  float* d_in_f32_[2] = {nullptr, nullptr};
  float* d_out_[2] = {nullptr, nullptr};
  void* h_in_[2] = {nullptr, nullptr};
  float* h_in_f32_[2] = {nullptr, nullptr};
  float* h_out_[2] = {nullptr, nullptr};
  void* l2_flush_ = nullptr;
  size_t in_slot_bytes_ = 0;
  std::vector<void*> dev_allocs_;

  // Compute lanes (round 2, opt-in with CLIPB200_LANES=2): consecutive micro-batches alternate between two compute
  // streams, each with its own activation workspace, so the HBM-bound kernels of one micro-batch (LayerNorm,
  // preprocessing, pooling) and the tail waves of its GEMMs run under the tensor-bound GEMMs of the other.  Measured
  // neutral under the power cap (engine.cu, AllocWorkspace), so one lane is the default.  `compute_` and the workspace
  // pointers above are the CURRENT lane's (BindLane); lane 0 is also the stream of the public event / synchronise /
  // flush entry points.
  struct Lane {
    __nv_bfloat16 *patches = nullptr, *h = nullptr, *qkv = nullptr, *mlpbuf = nullptr, *pooled = nullptr, *yh = nullptr,
                  *ymlp = nullptr, *vt = nullptr, *fv_stem_out = nullptr;
    float *x = nullptr, *y = nullptr, *proj_out = nullptr, *fv_xa = nullptr, *fv_xb = nullptr, *fv_tmp = nullptr,
          *fv_s = nullptr, *fv_gate = nullptr;
    int* row_map = nullptr;
    cudaStream_t stream = nullptr;
  };
  Lane lanes_[2];
  int n_lanes_ = 1, lane_ = 0;
  cudaEvent_t lane_fork_ = nullptr, lane_join_ = nullptr;
  void SaveLane(int k);
  void BindLane(int k);
  Status AllocActivations();
  cudaStream_t compute_ = nullptr, copy_in_ = nullptr, copy_out_ = nullptr;
  cudaStream_t resize_ = nullptr;   // photo path: resize kernels of group g run while copy_in_ uploads group g + 1
  cudaEvent_t in_ready_[2] = {nullptr, nullptr}, in_consumed_[2] = {nullptr, nullptr},
              out_ready_[2] = {nullptr, nullptr}, out_copied_[2] = {nullptr, nullptr};
  cudaEvent_t user_events_[16] = {};

  struct ProfPair {
    int cls;
    cudaEvent_t a, b;
  };
  std::vector<ProfPair> prof_pending_;
  std::vector<cudaEvent_t> prof_free_;
  clipb200_profile prof_acc_ = {};

  template <typename InT>
  Status RunPipelined(const InT* in, int64_t batch, size_t in_elems_per_item, float* out, bool device_buffers,
                      int mode);
};

}  // namespace clipb200
