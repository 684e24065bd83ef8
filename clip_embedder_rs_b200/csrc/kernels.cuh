// Non-GEMM kernels of the embedding engine (all HBM-bound except attention):
//   preprocess_patches_u8   u8 HWC image -> LUT-exact normalise -> bf16 patch matrix (fused im2col)   [vision.rs:235-259]
//   im2col_f32              f32 NCHW pixel_values (the ORT-style input) -> bf16 patch matrix           [vision.rs:105]
//   layernorm               fp32 residual stream -> bf16 (GEMM operand) or fp32 (ln_pre)
//   flash_attention         softmax(QK^T/sqrt(d)) V on mma.sync tensor cores, fp32 softmax
//   map_pool_attention      SigLIP attention-pool: one learned query x T keys per head
//   embed_tokens            token-embedding gather + positional add
//   l2_normalize            x / max(||x||, 1e-12)                                                      [pull_onnx.py:58-59,67-68]
//   similarity              dot -> fma(scale,bias) -> sigmoid | softmax                                [clip.rs:102-121,174-185]
//   weight conversion       fp32 -> bf16 (optional transpose / row padding)
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace clipb200 {

cudaError_t launch_preprocess_patches_u8(const uint8_t* img, int n, int S, int P, int Kp, const float* lut /*[3][256]*/,
                                         __nv_bfloat16* patches, cudaStream_t st);
cudaError_t launch_normalize_nchw_f32(const uint8_t* img, int n, int S, const float* lut, float* out, cudaStream_t st);
cudaError_t launch_im2col_f32(const float* nchw, int n, int S, int P, int Kp, __nv_bfloat16* patches, cudaStream_t st);

// rows: number of output rows; row_map (nullable): source row for each output row.
cudaError_t launch_layernorm(const float* x, const int* row_map, int rows, int D, const float* gamma,
                             const float* beta, float eps, __nv_bfloat16* out_bf16, float* out_f32, cudaStream_t st);

// qkv: [B*T, 3*H*hd] bf16 (q | k | v, each head-major); out: [B*T, H*hd] bf16
cudaError_t launch_flash_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int H, int hd,
                                   bool causal, cudaStream_t st);
cudaError_t flash_attention_configure_device();

// kv: [B*T, 2*H*hd] bf16 (k | v); q: [H*hd] f32 already scaled by hd^-0.5; out: [B, H*hd] bf16
cudaError_t launch_map_pool_attention(const __nv_bfloat16* kv, const float* q, __nv_bfloat16* out, int B, int T, int H,
                                      int hd, cudaStream_t st);

cudaError_t launch_embed_tokens(const int64_t* ids, int rows, int ctx, int D, int vocab, const float* tok_emb,
                                const float* pos_emb, float* x, int* err_flag, cudaStream_t st);
// row_map[b] = b*ctx + argmax_t ids[b,t] (first maximum) or b*ctx + ctx-1
cudaError_t launch_text_pool_rows(const int64_t* ids, int B, int ctx, bool argmax, int* row_map, cudaStream_t st);
// row_map[b] = b*stride + offset
cudaError_t launch_affine_rows(int B, int stride, int offset, int* row_map, cudaStream_t st);
// x[b*T + 0, :] = cls_row[:]   (class_embedding + positional_embedding[0], pre-added at load time)
cudaError_t launch_write_cls_rows(float* x, int B, int T, int D, const float* cls_row, cudaStream_t st);

cudaError_t launch_l2_normalize(const float* x, int rows, int D, float* out, cudaStream_t st);

// logits[i] = fma(dot(A[i,:], b), scale, bias); activation 0 = softmax over all N, 1 = sigmoid.
// scratch: >= 2 floats.
cudaError_t launch_similarity(const float* A, const float* b, int N, int D, float scale, float bias, int activation,
                              float* probs, float* scratch, cudaStream_t st);

// dst[r, c] (bf16, ld_out >= cols, zero padded) = src[r, c] (f32);  transpose: src is [cols, rows]
cudaError_t launch_convert_f32_bf16(const float* src, int rows, int cols, int ld_out, bool transpose,
                                    __nv_bfloat16* dst, cudaStream_t st);

}  // namespace clipb200
