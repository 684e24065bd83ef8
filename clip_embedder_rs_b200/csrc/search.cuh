// Multi-query corpus search kernels (search.cu) and the plain-function door to the tcgen05 GEMM (engine.cu) that the
// C-ABI corpus object uses.  Reference: `Clip::rank_images`, src/clip.rs:136-170.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace clipb200 {

// fp32 corpus rows [n, D] -> bf16 [n, 2D] = (hi | lo)
cudaError_t launch_split_corpus_rows(const float* rows, long long n, int D, __nv_bfloat16* dst, cudaStream_t st);
// fp32 queries [n, D] -> bf16 [n, 2D] = (hi | hi) and bf16 [n, D] = lo
cudaError_t launch_split_queries(const float* q, int n, int D, __nv_bfloat16* hi_hi, __nv_bfloat16* lo, cudaStream_t st);
// keys per query each of the two ping-pong buffers must hold
size_t search_scratch_keys(int N, int k);
// logits [n_queries, ld] fp32 -> top_index / top_prob [n_queries, k]; k <= 2048 and k <= N
cudaError_t launch_search_topk(const float* logits, long long ld, int n_queries, int N, int k, float scale, float bias,
                               int activation, float2* stats, unsigned long long* keys_a, unsigned long long* keys_b,
                               long long* top_index, float* top_prob, cudaStream_t st);

// out[M, ldc] (=, or += when accumulate) A[M, K] * W[N, K]^T on the engine's tcgen05 GEMM; all bf16 operands K-major,
// fp32 output.  Configures the kernels for the current device on first use.
cudaError_t gemm_bf16_f32out(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw, int M, int N,
                             int K, float* out, long long ldc, bool accumulate, cudaStream_t st);

}  // namespace clipb200
