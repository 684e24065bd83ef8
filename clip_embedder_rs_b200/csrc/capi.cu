// extern "C" boundary (include/clipb200.h).  Nothing throws across it; errors become status codes plus a
// thread-local message, the way `ort::Error` becomes `ClipError::Ort(String)` in the reference (src/error.rs:62-66).
#include <cuda_runtime.h>

#include <string.h>

#include <algorithm>
#include <new>
#include <string>

#include "../../include/clipb200.h"
#include "engine.h"
#include "kernels.cuh"
#include "search.cuh"

using clipb200::Engine;
using clipb200::Status;

struct clipb200_engine {
  Engine* impl;
};

static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
static int done(const Status& s) {
  if (s.ok()) return CLIPB200_OK;
  g_last_error = s.msg;
  return s.code;
}

// shared with pool.cu: sets this thread's message and returns the code
int clipb200_set_last_error(int code, const std::string& msg) { return fail(code, msg); }

#define API_GUARD_BEGIN try {
#define API_GUARD_END                                                        \
  }                                                                          \
  catch (const std::bad_alloc&) {                                            \
    return fail(CLIPB200_ERR_CUDA, "host allocation failed");                \
  }                                                                          \
  catch (const std::exception& ex) {                                         \
    return fail(CLIPB200_ERR_INVALID_ARG, std::string("exception: ") + ex.what()); \
  }                                                                          \
  catch (...) {                                                              \
    return fail(CLIPB200_ERR_INVALID_ARG, "unknown exception");              \
  }

extern "C" {

const char* clipb200_last_error(void) { return g_last_error.c_str(); }
const char* clipb200_version(void) { return "clipb200 0.1.0 (sm_100a)"; }

int clipb200_engine_create(const char* onnx_path, int cuda_device, const clipb200_opts* opts, clipb200_engine** out) {
  API_GUARD_BEGIN
  if (onnx_path == nullptr || out == nullptr) return fail(CLIPB200_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  Engine* e = nullptr;
  Status s = Engine::Create(onnx_path, cuda_device, opts, &e);
  if (!s.ok()) return done(s);
  clipb200_engine* h = new clipb200_engine();
  h->impl = e;
  *out = h;
  return CLIPB200_OK;
  API_GUARD_END
}

void clipb200_engine_destroy(clipb200_engine* e) {
  if (e == nullptr) return;
  try {
    delete e->impl;
    delete e;
  } catch (...) {
  }
}

// Parses the file exactly as clipb200_engine_create does, without touching a GPU, and describes what it found.
int clipb200_onnx_inspect(const char* onnx_path, char* json_out, size_t capacity) {
  API_GUARD_BEGIN
  if (onnx_path == nullptr || json_out == nullptr || capacity == 0) return fail(CLIPB200_ERR_INVALID_ARG, "null argument");
  clipb200::OnnxModel m;
  std::string err;
  if (!clipb200::load_onnx(onnx_path, &m, &err)) {
    const bool io = err.find("cannot open") != std::string::npos || err.find("cannot stat") != std::string::npos ||
                    err.find("cannot mmap") != std::string::npos || err.find("is empty") != std::string::npos;
    return fail(io ? CLIPB200_ERR_IO : CLIPB200_ERR_PARSE, err);
  }
  auto esc = [](const std::string& in) {
    std::string o;
    for (char c : in) {
      if (c == '"' || c == '\\') { o.push_back('\\'); o.push_back(c); }
      else if (static_cast<unsigned char>(c) < 0x20) o.push_back(' ');
      else o.push_back(c);
    }
    return o;
  };
  size_t bytes = 0;
  for (const auto& kv : m.initializers) bytes += kv.second.nbytes;
  const size_t n_file_initializers = m.initializers.size();
  // the same structural binding clipb200_engine_create performs for real exports (onnx_graph.h)
  std::vector<clipb200::GraphBinding> bindings;
  std::string graph_err;
  const bool graph_tried = clipb200::graph_needs_recognition(m);
  const bool graph_ok = graph_tried && clipb200::recognize_graph(&m, &graph_err, &bindings);
  // ... and for the one family bound by name, the FastViT trunk: its attention Linears come out of the graph
  bool fastvit_ok = false;
  std::string fastvit_err;
  if (graph_tried && !graph_ok && m.has("model.visual.trunk.stem.0.reparam_conv.weight"))
    fastvit_ok = clipb200::bind_fastvit_graph(&m, &fastvit_err);
  std::string j = "{\"inputs\": [";
  for (size_t i = 0; i < m.inputs.size(); ++i) j += std::string(i ? ", " : "") + "\"" + esc(m.inputs[i]) + "\"";
  j += "], \"outputs\": [";
  for (size_t i = 0; i < m.outputs.size(); ++i) j += std::string(i ? ", " : "") + "\"" + esc(m.outputs[i]) + "\"";
  j += "], \"opset\": " + std::to_string(m.opset) + ", \"num_initializers\": " + std::to_string(n_file_initializers) +
       ", \"initializer_bytes\": " + std::to_string(bytes) + ", \"num_nodes\": " + std::to_string(m.nodes.size()) +
       ", \"metadata\": {";
  bool first = true;
  for (const auto& kv : m.metadata) {
    j += std::string(first ? "" : ", ") + "\"" + esc(kv.first) + "\": \"" + esc(kv.second) + "\"";
    first = false;
  }
  j += "}, \"graph\": {\"attempted\": " + std::string(graph_tried ? "true" : "false") + ", \"recognized\": " +
       std::string(graph_ok ? "true" : "false") + ", \"fastvit_by_name\": " + std::string(fastvit_ok ? "true" : "false") +
       ", \"error\": \"" + esc(graph_err + (fastvit_err.empty() ? "" : "; " + fastvit_err)) + "\", \"bindings\": [";
  for (size_t i = 0; i < bindings.size(); ++i)
    j += std::string(i ? ", " : "") + "{\"name\": \"" + esc(bindings[i].canonical) + "\", \"source\": \"" +
         esc(bindings[i].source) + "\", \"transposed\": " + (bindings[i].transposed ? "true" : "false") + "}";
  j += "]}}";
  if (j.size() + 1 > capacity) return fail(CLIPB200_ERR_INVALID_ARG, "output buffer too small");
  memcpy(json_out, j.c_str(), j.size() + 1);
  return CLIPB200_OK;
  API_GUARD_END
}

// Parse-only: the fp32 value of one tensor exactly as clipb200_engine_create would bind it (after graph recognition,
// canonical [out, in] layout for Linear weights).
int clipb200_onnx_read_tensor(const char* onnx_path, const char* name, float* out, size_t capacity, int64_t* dims_out,
                              int* rank_out) {
  API_GUARD_BEGIN
  if (onnx_path == nullptr || name == nullptr || rank_out == nullptr) return fail(CLIPB200_ERR_INVALID_ARG, "null argument");
  clipb200::OnnxModel m;
  std::string err;
  if (!clipb200::load_onnx(onnx_path, &m, &err)) {
    const bool io = err.find("cannot open") != std::string::npos || err.find("cannot stat") != std::string::npos ||
                    err.find("cannot mmap") != std::string::npos || err.find("is empty") != std::string::npos;
    return fail(io ? CLIPB200_ERR_IO : CLIPB200_ERR_PARSE, err);
  }
  if (clipb200::graph_needs_recognition(m) && !clipb200::recognize_graph(&m, &err, nullptr)) {
    // not fatal: the tensor may still be present under its exported name
    std::string fv_err;
    if (m.has("model.visual.trunk.stem.0.reparam_conv.weight") && !clipb200::bind_fastvit_graph(&m, &fv_err)) err += "; " + fv_err;
  }
  const clipb200::OnnxTensor* t = m.find(name);
  if (t == nullptr) return fail(CLIPB200_ERR_UNSUPPORTED, std::string("tensor '") + name + "' not found" + (err.empty() ? "" : "; " + err));
  if (t->dims.size() > 8) return fail(CLIPB200_ERR_UNSUPPORTED, "tensor rank > 8");
  *rank_out = static_cast<int>(t->dims.size());
  if (dims_out != nullptr) for (size_t i = 0; i < t->dims.size(); ++i) dims_out[i] = t->dims[i];
  if (out == nullptr) return CLIPB200_OK;  // shape query
  if (static_cast<size_t>(t->numel()) > capacity) return fail(CLIPB200_ERR_INVALID_ARG, "output buffer too small");
  std::vector<float> v;
  if (!clipb200::tensor_to_f32(*t, &v)) return fail(CLIPB200_ERR_UNSUPPORTED, "tensor is not f32 / f16 / bf16");
  memcpy(out, v.data(), v.size() * 4);
  return CLIPB200_OK;
  API_GUARD_END
}

int clipb200_engine_num_inputs(const clipb200_engine* e) {
  return e == nullptr ? 0 : static_cast<int>(e->impl->input_names.size());
}
const char* clipb200_engine_input_name(const clipb200_engine* e, int i) {
  if (e == nullptr || i < 0 || i >= static_cast<int>(e->impl->input_names.size())) return nullptr;
  return e->impl->input_names[i].c_str();
}
int clipb200_engine_kind(const clipb200_engine* e) { return e == nullptr ? -1 : e->impl->kind; }
int64_t clipb200_engine_embed_dim(const clipb200_engine* e) { return e == nullptr ? 0 : e->impl->embed_dim; }
int64_t clipb200_engine_image_size(const clipb200_engine* e) { return e == nullptr ? 0 : e->impl->image_size; }
int64_t clipb200_engine_context_length(const clipb200_engine* e) { return e == nullptr ? 0 : e->impl->context_length; }
int64_t clipb200_engine_weight_bytes(const clipb200_engine* e) { return e == nullptr ? 0 : e->impl->weight_bytes; }
int64_t clipb200_engine_launch_count(const clipb200_engine* e) { return e == nullptr ? 0 : e->impl->launch_count; }

int clipb200_vision_embed_f32(clipb200_engine* e, const float* nchw, int64_t batch, float* out) {
  API_GUARD_BEGIN
  if (e == nullptr) return fail(CLIPB200_ERR_INVALID_ARG, "null engine");
  return done(e->impl->VisionEmbedF32(nchw, batch, out));
  API_GUARD_END
}

int clipb200_vision_embed_rgb8(clipb200_engine* e, const uint8_t* hwc, int64_t batch, int32_t width, int32_t height,
                               const clipb200_preproc* pp, float* out) {
  API_GUARD_BEGIN
  if (e == nullptr) return fail(CLIPB200_ERR_INVALID_ARG, "null engine");
  return done(e->impl->VisionEmbedRgb8(hwc, batch, width, height, pp, out, false));
  API_GUARD_END
}

int clipb200_preprocess_rgb8(clipb200_engine* e, const uint8_t* hwc, int64_t batch, int32_t width, int32_t height,
                             const clipb200_preproc* pp, float* out_nchw) {
  API_GUARD_BEGIN
  if (e == nullptr) return fail(CLIPB200_ERR_INVALID_ARG, "null engine");
  return done(e->impl->PreprocessRgb8(hwc, batch, width, height, pp, out_nchw));
  API_GUARD_END
}

int clipb200_vision_embed_rgb8_var(clipb200_engine* e, const uint8_t* const* images, const int32_t* widths,
                                   const int32_t* heights, int64_t batch, const clipb200_preproc* pp, float* out) {
  API_GUARD_BEGIN
  if (e == nullptr) return fail(CLIPB200_ERR_INVALID_ARG, "null engine");
  return done(e->impl->VisionEmbedRgb8Var(images, widths, heights, batch, pp, out));
  API_GUARD_END
}

int clipb200_resize_rgb8(clipb200_engine* e, const uint8_t* image, int32_t width, int32_t height,
                         const clipb200_preproc* pp, uint8_t* out) {
  API_GUARD_BEGIN
  if (e == nullptr) return fail(CLIPB200_ERR_INVALID_ARG, "null engine");
  return done(e->impl->ResizeRgb8(image, width, height, pp, out));
  API_GUARD_END
}

int clipb200_vision_embed_rgb8_device(clipb200_engine* e, const uint8_t* d_hwc, int64_t batch,
                                      const clipb200_preproc* pp, float* d_out) {
  API_GUARD_BEGIN
  if (e == nullptr) return fail(CLIPB200_ERR_INVALID_ARG, "null engine");
  const int s = static_cast<int>(e->impl->image_size);
  return done(e->impl->VisionEmbedRgb8(d_hwc, batch, s, s, pp, d_out, true));
  API_GUARD_END
}

int clipb200_text_embed(clipb200_engine* e, const int64_t* input_ids, const int64_t* /*attention_mask_or_null*/,
                        int64_t batch, int64_t ctx, float* out) {
  API_GUARD_BEGIN
  if (e == nullptr) return fail(CLIPB200_ERR_INVALID_ARG, "null engine");
  return done(e->impl->TextEmbed(input_ids, batch, ctx, out, false));
  API_GUARD_END
}

int clipb200_text_embed_device(clipb200_engine* e, const int64_t* d_input_ids, int64_t batch, int64_t ctx,
                               float* d_out) {
  API_GUARD_BEGIN
  if (e == nullptr) return fail(CLIPB200_ERR_INVALID_ARG, "null engine");
  return done(e->impl->TextEmbed(d_input_ids, batch, ctx, d_out, true));
  API_GUARD_END
}

int clipb200_similarity(int cuda_device, const float* A, const float* b, int64_t n, int64_t d, float scale,
                        float bias, int activation, float* probs) {
  API_GUARD_BEGIN
  if (A == nullptr || b == nullptr || probs == nullptr) return fail(CLIPB200_ERR_INVALID_ARG, "null buffer");
  if (n <= 0 || d <= 0 || n > 0x7fffffff || d > 0x7fffffff) return fail(CLIPB200_ERR_INVALID_ARG, "bad shape");
  cudaError_t ce = cudaSetDevice(cuda_device);
  if (ce != cudaSuccess) return fail(CLIPB200_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(ce));
  float *dA = nullptr, *db = nullptr, *dp = nullptr;
  const size_t abytes = static_cast<size_t>(n) * d * 4;
  int rc = CLIPB200_OK;
  do {
    if ((ce = cudaMalloc(&dA, abytes)) != cudaSuccess) break;
    if ((ce = cudaMalloc(&db, d * 4)) != cudaSuccess) break;
    if ((ce = cudaMalloc(&dp, n * 4 + 16)) != cudaSuccess) break;
    if ((ce = cudaMemcpy(dA, A, abytes, cudaMemcpyHostToDevice)) != cudaSuccess) break;
    if ((ce = cudaMemcpy(db, b, d * 4, cudaMemcpyHostToDevice)) != cudaSuccess) break;
    if ((ce = clipb200::launch_similarity(dA, db, static_cast<int>(n), static_cast<int>(d), scale, bias, activation, dp,
                                          nullptr, 0)) != cudaSuccess)
      break;
    if ((ce = cudaMemcpy(probs, dp, n * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
  } while (0);
  if (ce != cudaSuccess) rc = fail(CLIPB200_ERR_CUDA, std::string("similarity: ") + cudaGetErrorString(ce));
  cudaFree(dA);
  cudaFree(db);
  cudaFree(dp);
  return rc;
  API_GUARD_END
}

struct clipb200_corpus {
  int device = 0;
  int64_t dim = 0, capacity = 0, cap8 = 0, size = 0;
  float* rows = nullptr;            // [capacity, dim] fp32 (single-query GEMV path)
  __nv_bfloat16* split = nullptr;   // [cap8, 2 dim] bf16 (hi | lo): the GEMM operand of clipb200_corpus_search
  float* query = nullptr;
  float* probs = nullptr;
  // search scratch, grown on demand
  float* d_queries = nullptr;
  __nv_bfloat16 *q_hihi = nullptr, *q_lo = nullptr;
  float* logits = nullptr;
  float2* stats = nullptr;
  unsigned long long *keys_a = nullptr, *keys_b = nullptr;
  long long* d_index = nullptr;
  float* d_prob = nullptr;
  int64_t q_cap = 0, key_cap = 0, out_cap = 0;
  void free_search() {
    cudaFree(d_queries); cudaFree(q_hihi); cudaFree(q_lo); cudaFree(logits); cudaFree(stats);
    cudaFree(keys_a); cudaFree(keys_b); cudaFree(d_index); cudaFree(d_prob);
    d_queries = nullptr; q_hihi = q_lo = nullptr; logits = nullptr; stats = nullptr; keys_a = keys_b = nullptr;
    d_index = nullptr; d_prob = nullptr;
    q_cap = key_cap = out_cap = 0;
  }
};

int clipb200_corpus_create(int cuda_device, int64_t dim, int64_t capacity, clipb200_corpus** out) {
  API_GUARD_BEGIN
  if (out == nullptr || dim <= 0 || capacity <= 0 || dim > 0x7fffffff || capacity > 0x7ffffff0)
    return fail(CLIPB200_ERR_INVALID_ARG, "bad corpus shape");
  *out = nullptr;
  cudaError_t ce = cudaSetDevice(cuda_device);
  if (ce != cudaSuccess) return fail(CLIPB200_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(ce));
  clipb200_corpus* c = new clipb200_corpus();
  c->device = cuda_device;
  c->dim = dim;
  c->capacity = capacity;
  c->cap8 = (capacity + 7) & ~int64_t(7);
  const size_t split_bytes = static_cast<size_t>(c->cap8) * dim * 2 * 2;
  if ((ce = cudaMalloc(&c->rows, static_cast<size_t>(capacity) * dim * 4)) != cudaSuccess ||
      (ce = cudaMalloc(&c->split, split_bytes)) != cudaSuccess ||
      (ce = cudaMemset(c->split, 0, split_bytes)) != cudaSuccess ||  // rows beyond `size` take part in the last N tile
      (ce = cudaMalloc(&c->query, dim * 4)) != cudaSuccess || (ce = cudaMalloc(&c->probs, capacity * 4 + 16)) != cudaSuccess) {
    cudaFree(c->rows); cudaFree(c->split); cudaFree(c->query); cudaFree(c->probs);
    delete c;
    return fail(CLIPB200_ERR_CUDA, std::string("corpus allocation: ") + cudaGetErrorString(ce));
  }
  *out = c;
  return CLIPB200_OK;
  API_GUARD_END
}
void clipb200_corpus_destroy(clipb200_corpus* c) {
  if (c == nullptr) return;
  cudaSetDevice(c->device);
  cudaFree(c->rows); cudaFree(c->split); cudaFree(c->query); cudaFree(c->probs);
  c->free_search();
  delete c;
}
int clipb200_corpus_append(clipb200_corpus* c, const float* rows, int64_t n) {
  API_GUARD_BEGIN
  if (c == nullptr || rows == nullptr || n < 0) return fail(CLIPB200_ERR_INVALID_ARG, "null argument");
  if (c->size + n > c->capacity) return fail(CLIPB200_ERR_INVALID_ARG, "corpus capacity exceeded");
  cudaError_t ce = cudaSetDevice(c->device);
  float* dst = c->rows + c->size * c->dim;
  if (ce == cudaSuccess) ce = cudaMemcpy(dst, rows, static_cast<size_t>(n) * c->dim * 4, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess)
    ce = clipb200::launch_split_corpus_rows(dst, n, static_cast<int>(c->dim), c->split + c->size * c->dim * 2, nullptr);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(nullptr);
  if (ce != cudaSuccess) return fail(CLIPB200_ERR_CUDA, std::string("corpus append: ") + cudaGetErrorString(ce));
  c->size += n;
  return CLIPB200_OK;
  API_GUARD_END
}
int64_t clipb200_corpus_size(const clipb200_corpus* c) { return c == nullptr ? 0 : c->size; }
int clipb200_corpus_rank(clipb200_corpus* c, const float* query, float scale, float bias, int activation, float* probs) {
  API_GUARD_BEGIN
  if (c == nullptr || query == nullptr || probs == nullptr) return fail(CLIPB200_ERR_INVALID_ARG, "null argument");
  if (c->size == 0) return fail(CLIPB200_ERR_INVALID_ARG, "Empty batch");
  cudaError_t ce = cudaSetDevice(c->device);
  if (ce == cudaSuccess) ce = cudaMemcpy(c->query, query, c->dim * 4, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess)
    ce = clipb200::launch_similarity(c->rows, c->query, static_cast<int>(c->size), static_cast<int>(c->dim), scale, bias,
                                     activation, c->probs, nullptr, 0);
  if (ce == cudaSuccess) ce = cudaMemcpy(probs, c->probs, c->size * 4, cudaMemcpyDeviceToHost);
  if (ce != cudaSuccess) return fail(CLIPB200_ERR_CUDA, std::string("corpus rank: ") + cudaGetErrorString(ce));
  return CLIPB200_OK;
  API_GUARD_END
}

// `rank_images` (src/clip.rs:136-170) for n_queries text embeddings at once: logits [Q, n] on the tcgen05 GEMM
// (split-bf16 operands, fp32-grade), softmax-over-corpus / sigmoid and the top-k selection on the GPU.
int clipb200_corpus_search(clipb200_corpus* c, const float* queries, int64_t n_queries, int64_t k, float scale, float bias,
                           int activation, int64_t* top_index, float* top_prob) {
  API_GUARD_BEGIN
  if (c == nullptr || queries == nullptr || top_index == nullptr || top_prob == nullptr)
    return fail(CLIPB200_ERR_INVALID_ARG, "null argument");
  if (c->size == 0 || n_queries <= 0) return fail(CLIPB200_ERR_INVALID_ARG, "Empty batch");
  if (k <= 0 || k > 2048 || k > c->size) return fail(CLIPB200_ERR_INVALID_ARG, "k must be in [1, min(2048, corpus size)]");
  if (c->dim % 8 != 0) return fail(CLIPB200_ERR_UNSUPPORTED, "corpus search needs an embedding width that is a multiple of 8");
  if (activation < 0 || activation > 2) return fail(CLIPB200_ERR_INVALID_ARG, "bad activation");
  cudaError_t ce = cudaSetDevice(c->device);
  if (ce != cudaSuccess) return fail(CLIPB200_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(ce));
  const int D = static_cast<int>(c->dim);
  const int N = static_cast<int>(c->size);
  const int N8 = (N + 7) & ~7;                   // GEMM column count; rows [N, N8) of `split` are zero
  const long long ld = N8;
  constexpr int64_t kQueryBlock = 128;           // one GEMM row tile per pass over the corpus
  const size_t keys = clipb200::search_scratch_keys(N, static_cast<int>(k));
  if (c->q_cap < kQueryBlock || c->key_cap < static_cast<int64_t>(keys) || c->out_cap < k) {
    c->free_search();
    const size_t Q = kQueryBlock;
    if ((ce = cudaMalloc(&c->d_queries, Q * D * 4)) != cudaSuccess || (ce = cudaMalloc(&c->q_hihi, Q * D * 2 * 2)) != cudaSuccess ||
        (ce = cudaMalloc(&c->q_lo, Q * D * 2)) != cudaSuccess ||
        (ce = cudaMalloc(&c->logits, Q * static_cast<size_t>(c->cap8) * 4)) != cudaSuccess ||
        (ce = cudaMalloc(&c->stats, Q * sizeof(float2))) != cudaSuccess ||
        (ce = cudaMalloc(&c->keys_a, Q * keys * 8 + 8)) != cudaSuccess || (ce = cudaMalloc(&c->keys_b, Q * keys * 8 + 8)) != cudaSuccess ||
        (ce = cudaMalloc(&c->d_index, Q * k * 8)) != cudaSuccess || (ce = cudaMalloc(&c->d_prob, Q * k * 4)) != cudaSuccess) {
      c->free_search();
      return fail(CLIPB200_ERR_CUDA, std::string("corpus search scratch: ") + cudaGetErrorString(ce));
    }
    c->q_cap = kQueryBlock;
    c->key_cap = static_cast<int64_t>(keys);
    c->out_cap = k;
  }
  for (int64_t q0 = 0; q0 < n_queries && ce == cudaSuccess; q0 += kQueryBlock) {
    const int nq = static_cast<int>(std::min<int64_t>(kQueryBlock, n_queries - q0));
    ce = cudaMemcpyAsync(c->d_queries, queries + q0 * D, static_cast<size_t>(nq) * D * 4, cudaMemcpyHostToDevice, nullptr);
    if (ce == cudaSuccess) ce = clipb200::launch_split_queries(c->d_queries, nq, D, c->q_hihi, c->q_lo, nullptr);
    // logits = q_hi.(c_hi + c_lo) over K = 2D, then += q_lo.c_hi over the first D columns of the same corpus rows
    if (ce == cudaSuccess)
      ce = clipb200::gemm_bf16_f32out(c->q_hihi, 2 * D, c->split, 2 * D, nq, N8, 2 * D, c->logits, ld, false, nullptr);
    if (ce == cudaSuccess)
      ce = clipb200::gemm_bf16_f32out(c->q_lo, D, c->split, 2 * D, nq, N8, D, c->logits, ld, true, nullptr);
    if (ce == cudaSuccess)
      ce = clipb200::launch_search_topk(c->logits, ld, nq, N, static_cast<int>(k), scale, bias, activation, c->stats,
                                        c->keys_a, c->keys_b, c->d_index, c->d_prob, nullptr);
    if (ce == cudaSuccess)
      ce = cudaMemcpyAsync(top_index + q0 * k, c->d_index, static_cast<size_t>(nq) * k * 8, cudaMemcpyDeviceToHost, nullptr);
    if (ce == cudaSuccess)
      ce = cudaMemcpyAsync(top_prob + q0 * k, c->d_prob, static_cast<size_t>(nq) * k * 4, cudaMemcpyDeviceToHost, nullptr);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(nullptr);
  }
  if (ce != cudaSuccess) return fail(CLIPB200_ERR_CUDA, std::string("corpus search: ") + cudaGetErrorString(ce));
  return CLIPB200_OK;
  API_GUARD_END
}

void* clipb200_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void clipb200_host_free(void* p) {
  if (p != nullptr) cudaFreeHost(p);
}
void* clipb200_device_alloc(int cuda_device, size_t bytes) {
  void* p = nullptr;
  if (cudaSetDevice(cuda_device) != cudaSuccess) return nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
  return p;
}
void clipb200_device_free(int cuda_device, void* p) {
  if (p == nullptr) return;
  cudaSetDevice(cuda_device);
  cudaFree(p);
}
int clipb200_memcpy_h2d(int cuda_device, void* dst, const void* src, size_t bytes) {
  cudaError_t e = cudaSetDevice(cuda_device);
  if (e == cudaSuccess) e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
  return e == cudaSuccess ? CLIPB200_OK : fail(CLIPB200_ERR_CUDA, cudaGetErrorString(e));
}
int clipb200_memcpy_d2h(int cuda_device, void* dst, const void* src, size_t bytes) {
  cudaError_t e = cudaSetDevice(cuda_device);
  if (e == cudaSuccess) e = cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? CLIPB200_OK : fail(CLIPB200_ERR_CUDA, cudaGetErrorString(e));
}
int clipb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int clipb200_engine_record_event(clipb200_engine* e, int slot) {
  API_GUARD_BEGIN
  if (e == nullptr) return fail(CLIPB200_ERR_INVALID_ARG, "null engine");
  return done(e->impl->RecordEvent(slot));
  API_GUARD_END
}
int clipb200_engine_elapsed_ms(clipb200_engine* e, int slot_start, int slot_end, double* ms) {
  API_GUARD_BEGIN
  if (e == nullptr) return fail(CLIPB200_ERR_INVALID_ARG, "null engine");
  return done(e->impl->ElapsedMs(slot_start, slot_end, ms));
  API_GUARD_END
}
int clipb200_engine_synchronize(clipb200_engine* e) {
  API_GUARD_BEGIN
  if (e == nullptr) return fail(CLIPB200_ERR_INVALID_ARG, "null engine");
  return done(e->impl->Synchronize());
  API_GUARD_END
}
int clipb200_engine_profile(clipb200_engine* e, clipb200_profile* out, int reset) {
  API_GUARD_BEGIN
  if (e == nullptr) return fail(CLIPB200_ERR_INVALID_ARG, "null engine");
  return done(e->impl->ReadProfile(out, reset != 0));
  API_GUARD_END
}
int clipb200_engine_flush_l2(clipb200_engine* e) {
  API_GUARD_BEGIN
  if (e == nullptr) return fail(CLIPB200_ERR_INVALID_ARG, "null engine");
  return done(e->impl->FlushL2());
  API_GUARD_END
}

}  // extern "C"
