/* clipb200 — C ABI of the B200-native CLIP / SigLIP embedding engine.
 *
 * This is the drop-in seam for `open_clip_inference` (RuurdBijlsma/clip-embedder-rs): every entry point
 * replaces one use of the `ort` crate in the reference.  Citations are paths under /root/reference.
 *
 *   reference (Rust, ort)                                           this library
 *   ------------------------------------------------------------    -----------------------------------------
 *   OnnxSession::new -> Session::builder()...commit_from_file       clipb200_engine_create
 *        (src/onnx.rs:14-29)
 *   Drop for ort::Session                                           clipb200_engine_destroy
 *   From<ort::Error> for ClipError::Ort(String) (src/error.rs:62)   status codes + clipb200_last_error
 *   session.inputs().iter().any(|i| i.name()==p)                    clipb200_engine_num_inputs / _input_name
 *        (src/onnx.rs:32-46; probed names src/vision.rs:73-75, src/text.rs:87-90)
 *   session.run(inputs![pixel_values => f32[B,3,S,S]])              clipb200_vision_embed_f32
 *        (src/vision.rs:105-113)
 *   preprocess_batch + session.run  (src/vision.rs:102-135,235-259) clipb200_vision_embed_rgb8  (GPU preprocessing)
 *   session.run(inputs![input_ids => i64[B,ctx] (, mask)])          clipb200_text_embed
 *        (src/text.rs:153-166)
 *   dot -> mul_add(scale,bias) -> softmax | sigmoid                 clipb200_similarity
 *        (src/clip.rs:102-121, 144-163, 174-185)
 *
 * Conventions: plain pointers and sizes only; inputs are borrowed for the duration of the call; the caller
 * allocates outputs; a handle is not re-entrant (the Rust side keeps it behind its RwLock write guard exactly
 * like `session.run`, src/vision.rs:107) but distinct handles may be used concurrently from different threads,
 * on the same or different GPUs.  There is NO CPU fallback: without an sm_100 device every call fails with
 * CLIPB200_ERR_CUDA.  Nothing throws or aborts across this boundary.
 */
#ifndef CLIPB200_H_
#define CLIPB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLIPB200_OK 0
#define CLIPB200_ERR_INVALID_ARG 1  /* null pointer, empty batch (src/vision.rs:121-123), wrong ctx ... */
#define CLIPB200_ERR_IO 2           /* missing / unreadable .onnx or .onnx.data                          */
#define CLIPB200_ERR_PARSE 3        /* not an ONNX ModelProto                                            */
#define CLIPB200_ERR_CUDA 4         /* no sm_100 device, allocation or launch failure                    */
#define CLIPB200_ERR_UNSUPPORTED 5  /* graph / architecture the engine cannot bind                       */

#define CLIPB200_KIND_VISION 0
#define CLIPB200_KIND_TEXT 1

#define CLIPB200_ACT_SOFTMAX 0
#define CLIPB200_ACT_SIGMOID 1
#define CLIPB200_ACT_NONE 2 /* raw logits: Clip::compare (src/clip.rs:81-90) */

typedef struct clipb200_engine clipb200_engine;

typedef struct clipb200_opts {
  int32_t micro_batch;   /* images / texts per internal pipeline step; 0 = automatic                     */
  int32_t profile;       /* 1 = time every kernel class with CUDA events (see clipb200_engine_profile)   */
  int32_t reserved[6];
} clipb200_opts;

/* preprocess_cfg of open_clip_config.json (src/config.rs:49-57) */
typedef struct clipb200_preproc {
  float mean[3];
  float std[3];
  int32_t interpolation; /* 0 bicubic (CatmullRom), 1 bilinear, 2 nearest  (src/vision.rs:176-180)        */
  int32_t resize_mode;   /* 0 shortest (centre crop), 1 squash              (src/vision.rs:184-192)       */
} clipb200_preproc;

/* Per-kernel-class device time accumulated since the last reset (profile = 1). */
#define CLIPB200_PROF_CLASSES 8
typedef struct clipb200_profile {
  double ms[CLIPB200_PROF_CLASSES];        /* 0 gemm, 1 attention, 2 layernorm, 3 preprocess, 4 pool/misc,  */
  int64_t launches[CLIPB200_PROF_CLASSES]; /* 5 h2d copy, 6 d2h copy, 7 depthwise conv (FastViT stages)     */
  double gemm_flops;                       /* algorithmic 2*M*N*K of the timed GEMM launches (no padding)   */
  double conv_bytes;                       /* algorithmic bytes (input + output once) of the timed depthwise convs */
} clipb200_profile;

/* ---- lifetime (src/onnx.rs:14-29) ------------------------------------------------------------------- */
int clipb200_engine_create(const char* onnx_path, int cuda_device, const clipb200_opts* opts_or_null,
                           clipb200_engine** out);
void clipb200_engine_destroy(clipb200_engine* e);
const char* clipb200_last_error(void); /* thread-local, valid until the next failing call on this thread */
const char* clipb200_version(void);

/* Parse-only probe (no GPU needed): writes a JSON description (inputs, outputs, opset, initializer count/bytes,
 * metadata) of what clipb200_engine_create would load.  Same IO / PARSE status codes as create. */
int clipb200_onnx_inspect(const char* onnx_path, char* json_out, size_t capacity);

/* Parse-only probe (no GPU needed): copies one tensor as fp32 exactly as clipb200_engine_create would bind it.
 * For files that carry an executable graph (what `torch.onnx.export` writes, pull_onnx.py:169-181) the tensors are
 * bound from the graph structure, not from initializer names, and are reported under the open_clip / timm
 * parameter names with Linear weights in [out, in] layout; `clipb200_onnx_inspect` lists the bindings under
 * "graph".  `out` may be NULL to query the shape only; `dims_out` must have room for 8 entries. */
int clipb200_onnx_read_tensor(const char* onnx_path, const char* name, float* out, size_t capacity_elems,
                              int64_t* dims_out, int* rank_out);

/* ---- introspection (src/onnx.rs:32-46) ---------------------------------------------------------------- */
int clipb200_engine_num_inputs(const clipb200_engine* e);
const char* clipb200_engine_input_name(const clipb200_engine* e, int i);
int clipb200_engine_kind(const clipb200_engine* e);
int64_t clipb200_engine_embed_dim(const clipb200_engine* e);
int64_t clipb200_engine_image_size(const clipb200_engine* e);      /* vision only, else 0 */
int64_t clipb200_engine_context_length(const clipb200_engine* e);  /* text only, else 0   */
int64_t clipb200_engine_weight_bytes(const clipb200_engine* e);    /* HBM held by weights */

/* ---- run: host buffers in, host buffers out ----------------------------------------------------------- */
/* ORT-equivalent: pixel_values f32 [B,3,S,S] -> out f32 [B,D] (rows L2-normalised). src/vision.rs:105-113 */
int clipb200_vision_embed_f32(clipb200_engine* e, const float* nchw, int64_t batch, float* out);
/* Fast path: packed RGB8 HWC images already at the model resolution, [B,S,S,3]; normalisation
 * (src/vision.rs:235-259) runs on the GPU.  Other sizes: use clipb200_vision_embed_rgb8_var. */
int clipb200_vision_embed_rgb8(clipb200_engine* e, const uint8_t* hwc, int64_t batch, int32_t width, int32_t height,
                               const clipb200_preproc* pp, float* out);
/* Images of arbitrary size (one pointer / width / height per image, packed RGB8 HWC): the resize of
 * `resize_with_fast_image_resize` (src/vision.rs:164-198: antialiased CatmullRom / bilinear convolution or nearest,
 * f64 centre-crop box unless resize_mode is squash) runs on the GPU, then normalise + tower as above. */
int clipb200_vision_embed_rgb8_var(clipb200_engine* e, const uint8_t* const* images, const int32_t* widths,
                                   const int32_t* heights, int64_t batch, const clipb200_preproc* pp, float* out);
/* Only the resize: one image in, S x S x 3 RGB8 out (what src/vision.rs:197 returns). */
int clipb200_resize_rgb8(clipb200_engine* e, const uint8_t* image, int32_t width, int32_t height,
                         const clipb200_preproc* pp, uint8_t* out);
/* The reference's public `preprocess_batch` (src/vision.rs:120-135): same inputs, out f32 [B,3,S,S]; bit-exact. */
int clipb200_preprocess_rgb8(clipb200_engine* e, const uint8_t* hwc, int64_t batch, int32_t width, int32_t height,
                             const clipb200_preproc* pp, float* out_nchw);
/* input_ids i64 [B,ctx].  Graphs written by pull_onnx.py have no attention_mask input (padding is handled inside the
 * graph: argmax / last-token pooling, src/text.rs:156-161 then passes none) and the pointer is not read; a graph that
 * DOES declare an attention_mask input is refused by clipb200_engine_create with CLIPB200_ERR_UNSUPPORTED. */
int clipb200_text_embed(clipb200_engine* e, const int64_t* input_ids, const int64_t* attention_mask_or_null,
                        int64_t batch, int64_t ctx, float* out);
/* probs[i] = act( fma(dot(A[i,:], b), scale, bias) ), softmax taken over all n rows.  src/clip.rs:102-121 */
int clipb200_similarity(int cuda_device, const float* A, const float* b, int64_t n, int64_t d, float scale,
                        float bias, int activation, float* probs);

/* ---- corpus search tail (src/clip.rs:136-170 `rank_images` at scale) -----------------------------
 * An embedding matrix [n, dim] kept resident in HBM.  clipb200_corpus_rank: one query, one similarity pass (HBM-bound
 * GEMV with the fused logit scale / bias / softmax-over-corpus or sigmoid), all n probabilities returned and the
 * stable descending sort left to the host.  clipb200_corpus_search (below): many queries, GEMM + GPU top-k. */
typedef struct clipb200_corpus clipb200_corpus;
int clipb200_corpus_create(int cuda_device, int64_t dim, int64_t capacity, clipb200_corpus** out);
void clipb200_corpus_destroy(clipb200_corpus* c);
int clipb200_corpus_append(clipb200_corpus* c, const float* rows, int64_t n);   /* host rows [n, dim] */
int64_t clipb200_corpus_size(const clipb200_corpus* c);
int clipb200_corpus_rank(clipb200_corpus* c, const float* query, float scale, float bias, int activation,
                         float* probs /* host [size] */);

/* `rank_images` for n_queries query embeddings at once (queries [n_queries, dim], host): logits [n_queries, size] in ONE
 * pass over the corpus on the tcgen05 GEMM (fp32 values split into bf16 hi + lo parts: fp32-grade dot products),
 * mul_add(scale, bias), softmax over the whole corpus / sigmoid / raw, and the per-query top-k on the GPU.  Returns, per
 * query, the first k entries of the reference's stable descending sort (src/clip.rs:167: ties keep the lower index
 * first): top_index / top_prob are host [n_queries, k].  1 <= k <= min(2048, size); dim must be a multiple of 8. */
int clipb200_corpus_search(clipb200_corpus* c, const float* queries, int64_t n_queries, int64_t k, float scale,
                           float bias, int activation, int64_t* top_index, float* top_prob);

/* ---- in-process multi-GPU pool: `duplicate()` (src/vision.rs:87-91, src/text.rs:104-108, src/clip.rs:69-73) for the
 * GPUs of one box.  One engine replica and one host thread per device; a call splits the batch into contiguous row
 * ranges (the first batch % n replicas take one extra row), every replica runs its own H2D / compute / D2H pipeline
 * and writes its embeddings straight into its rows of `out`.  No collective on the data path.  `devices` NULL or
 * n_devices <= 0 = all visible devices.  Results are bit-identical to a single engine's (same kernels, rows are
 * independent).  A pool handle is not re-entrant, like an engine handle. */
typedef struct clipb200_pool clipb200_pool;
int clipb200_pool_create(const char* onnx_path, const int32_t* devices, int32_t n_devices,
                         const clipb200_opts* opts_or_null, clipb200_pool** out);
void clipb200_pool_destroy(clipb200_pool* p);
int clipb200_pool_size(const clipb200_pool* p);                 /* number of replicas */
int clipb200_pool_device(const clipb200_pool* p, int replica);  /* CUDA device index of a replica, -1 if out of range */
int clipb200_pool_kind(const clipb200_pool* p);
int64_t clipb200_pool_embed_dim(const clipb200_pool* p);
int64_t clipb200_pool_image_size(const clipb200_pool* p);
int64_t clipb200_pool_context_length(const clipb200_pool* p);
int clipb200_pool_num_inputs(const clipb200_pool* p);           /* src/onnx.rs:32-46 on replica 0 */
const char* clipb200_pool_input_name(const clipb200_pool* p, int i);
int64_t clipb200_pool_launch_count(const clipb200_pool* p);     /* kernels launched by all replicas so far */
int clipb200_pool_vision_embed_rgb8(clipb200_pool* p, const uint8_t* hwc, int64_t batch, int32_t width, int32_t height,
                                    const clipb200_preproc* pp, float* out);
int clipb200_pool_vision_embed_rgb8_var(clipb200_pool* p, const uint8_t* const* images, const int32_t* widths,
                                        const int32_t* heights, int64_t batch, const clipb200_preproc* pp, float* out);
int clipb200_pool_text_embed(clipb200_pool* p, const int64_t* input_ids, const int64_t* attention_mask_or_null,
                             int64_t batch, int64_t ctx, float* out);

/* ---- run: device-resident buffers (benchmark "value": inputs already in HBM) --------------------------- */
int clipb200_vision_embed_rgb8_device(clipb200_engine* e, const uint8_t* d_hwc, int64_t batch,
                                      const clipb200_preproc* pp, float* d_out);
int clipb200_text_embed_device(clipb200_engine* e, const int64_t* d_input_ids, int64_t batch, int64_t ctx,
                               float* d_out);

/* ---- memory / timing helpers for callers without their own CUDA binding -------------------------------- */
void* clipb200_host_alloc(size_t bytes);                 /* pinned host memory (cudaHostAlloc) */
void clipb200_host_free(void* p);
void* clipb200_device_alloc(int cuda_device, size_t bytes);
void clipb200_device_free(int cuda_device, void* p);
int clipb200_memcpy_h2d(int cuda_device, void* dst, const void* src, size_t bytes);
int clipb200_memcpy_d2h(int cuda_device, void* dst, const void* src, size_t bytes);
int clipb200_device_count(void);
/* CUDA events on the engine's compute stream: record into slot 0..15, read elapsed ms between two slots. */
int clipb200_engine_record_event(clipb200_engine* e, int slot);
int clipb200_engine_elapsed_ms(clipb200_engine* e, int slot_start, int slot_end, double* ms);
int clipb200_engine_synchronize(clipb200_engine* e);
int clipb200_engine_profile(clipb200_engine* e, clipb200_profile* out, int reset);
int64_t clipb200_engine_launch_count(const clipb200_engine* e); /* kernels launched by this handle so far */
/* Writes > L2-size bytes on the engine's stream so the next step starts with a cold L2. */
int clipb200_engine_flush_l2(clipb200_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* CLIPB200_H_ */
