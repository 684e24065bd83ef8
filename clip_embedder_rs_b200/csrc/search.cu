// Multi-query corpus search: `Clip::rank_images` (reference src/clip.rs:136-170) for many text queries against an
// HBM-resident embedding matrix, with the ranking done on the GPU.
//
//   logits[q, i] = <query_q, corpus_i>                      one pass over the corpus on the tcgen05 GEMM
//   score[q, i]  = fma(logit, scale, bias)                  clip.rs:151-152 (f32::mul_add)
//   prob[q, i]   = softmax over the whole corpus | sigmoid  clip.rs:155-163
//   top-k        = stable descending sort by prob, first k  clip.rs:167 (ties keep the lower index first)
//
// Precision: the reference's dot product is fp32.  The tensor cores take bf16, so every fp32 value is split into
// hi = bf16(x) and lo = bf16(x - hi) and the product is evaluated as hi*hi + lo*hi + hi*lo with fp32 accumulation
// (the dropped lo*lo term is below 2^-16 relative): fp32-grade logits out of two GEMM launches over a [n, 2D] bf16
// copy of the corpus that occupies the same bytes as the fp32 rows.
// Top-k: 64-bit keys (order-preserving bits of the score << 32 | ~index) are sorted 4096 at a time in shared memory
// (bitonic network) and the best k of every chunk survive to the next round until one chunk is left: exact, stable
// in the reference's sense, and no library call.
#include "search.cuh"

#include <math.h>

namespace clipb200 {

__device__ __forceinline__ float warp_max_f(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// fp32 [rows, D] -> bf16 [rows, ld]: columns [0, D) = hi, [D, 2D) = lo (MODE 0, corpus rows);
//                                     columns [0, D) = hi, [D, 2D) = hi (MODE 1, first query operand);
//                                     columns [0, D) = lo             (MODE 2, second query operand, ld = D)
template <int MODE>
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ src, long long rows, int D, int ld, __nv_bfloat16* __restrict__ dst) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= rows * D) return;
  const long long r = idx / D;
  const int c = static_cast<int>(idx - r * D);
  const float x = src[idx];
  const __nv_bfloat16 hi = __float2bfloat16(x);
  const __nv_bfloat16 lo = __float2bfloat16(x - __bfloat162float(hi));
  __nv_bfloat16* o = dst + r * ld;
  if (MODE == 0) { o[c] = hi; o[D + c] = lo; }
  if (MODE == 1) { o[c] = hi; o[D + c] = hi; }
  if (MODE == 2) { o[c] = lo; }
}

cudaError_t launch_split_corpus_rows(const float* rows, long long n, int D, __nv_bfloat16* dst, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  split_bf16_kernel<0><<<static_cast<unsigned>((n * D + 255) / 256), 256, 0, st>>>(rows, n, D, 2 * D, dst);
  return cudaGetLastError();
}
cudaError_t launch_split_queries(const float* q, int n, int D, __nv_bfloat16* hi_hi, __nv_bfloat16* lo, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const unsigned blocks = static_cast<unsigned>((static_cast<long long>(n) * D + 255) / 256);
  split_bf16_kernel<1><<<blocks, 256, 0, st>>>(q, n, D, 2 * D, hi_hi);
  split_bf16_kernel<2><<<blocks, 256, 0, st>>>(q, n, D, D, lo);
  return cudaGetLastError();
}

// one block per query: max and sum(exp(score - max)) of score = fma(logit, scale, bias) over the corpus
__global__ void __launch_bounds__(1024)
search_stats_kernel(const float* __restrict__ logits, long long ld, int N, float scale, float bias,
                    float2* __restrict__ stats) {
  __shared__ float red[32];
  __shared__ float bc;
  const float* row = logits + blockIdx.x * ld;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float m = -INFINITY;
  for (int i = tid; i < N; i += 1024) m = fmaxf(m, fmaf(row[i], scale, bias));
  m = warp_max_f(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (warp == 0) {
    float v = warp_max_f(red[lane]);
    if (lane == 0) bc = v;
  }
  __syncthreads();
  const float gmax = bc;
  float s = 0.f;
  for (int i = tid; i < N; i += 1024) s += expf(fmaf(row[i], scale, bias) - gmax);
  s = warp_sum_f(s);
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (warp == 0) {
    float v = warp_sum_f(red[lane]);
    if (lane == 0) stats[blockIdx.x] = make_float2(gmax, v);
  }
}

constexpr int kChunk = 4096;      // keys sorted per block
constexpr int kSortThreads = 512;

__device__ __forceinline__ unsigned long long make_key(float score, unsigned idx) {
  unsigned b = __float_as_uint(score);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // unsigned order == float order
  return (static_cast<unsigned long long>(b) << 32) | (0xFFFFFFFFu - idx);  // ties: the lower index is the larger key
}
__device__ __forceinline__ float key_score(unsigned long long k) {
  unsigned b = static_cast<unsigned>(k >> 32);
  b = (b & 0x80000000u) ? (b & 0x7FFFFFFFu) : ~b;
  return __uint_as_float(b);
}

// Sorts one chunk of one query descending and keeps its best `keep` keys.
//   FIRST: keys come from the logits row (chunk c covers corpus rows [c * kChunk, ...)); else from the previous round.
//   LAST:  the surviving keys are decoded into (index, probability).
template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(kSortThreads)
topk_round_kernel(const float* __restrict__ logits, long long ld, const unsigned long long* __restrict__ in,
                  long long in_per_query, int n_in, float scale, float bias, int keep,
                  unsigned long long* __restrict__ out, long long out_per_query, const float2* __restrict__ stats,
                  int activation, long long* __restrict__ top_index, float* __restrict__ top_prob, int k_out) {
  __shared__ unsigned long long s[kChunk];
  const int q = blockIdx.y, c = blockIdx.x, tid = threadIdx.x;
  const int base = c * kChunk;
  for (int i = tid; i < kChunk; i += kSortThreads) {
    const int g = base + i;
    unsigned long long key = 0ull;  // below every real key
    if (g < n_in) {
      if (FIRST) {
        const float sc = fmaf(logits[q * ld + g], scale, bias);
        key = make_key(sc == sc ? sc : -INFINITY, static_cast<unsigned>(g));  // NaN ranks last
      } else {
        key = in[q * in_per_query + g];
      }
    }
    s[i] = key;
  }
  __syncthreads();
  for (int k = 2; k <= kChunk; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < kChunk; i += kSortThreads) {
        const int p = i ^ j;
        if (p > i) {
          const unsigned long long a = s[i], b = s[p];
          const bool desc = (i & k) == 0;
          if (desc ? (a < b) : (a > b)) { s[i] = b; s[p] = a; }
        }
      }
      __syncthreads();
    }
  }
  if (!LAST) {
    for (int i = tid; i < keep; i += kSortThreads) out[q * out_per_query + static_cast<long long>(c) * keep + i] = s[i];
  } else {
    const float2 st = stats[q];
    for (int i = tid; i < k_out; i += kSortThreads) {
      const unsigned long long key = s[i];
      const float sc = key_score(key);
      float p = sc;                                            // CLIPB200_ACT_NONE: the raw logit
      if (activation == 0) p = expf(sc - st.x) / st.y;         // softmax over the whole corpus (clip.rs:155-159)
      else if (activation == 1) p = 1.0f / (1.0f + expf(-sc)); // clip.rs:160-162
      top_index[static_cast<long long>(q) * k_out + i] = static_cast<long long>(0xFFFFFFFFu - static_cast<unsigned>(key));
      top_prob[static_cast<long long>(q) * k_out + i] = p;
    }
  }
}

size_t search_scratch_keys(int N, int k) {
  const long long chunks = (static_cast<long long>(N) + kChunk - 1) / kChunk;
  return static_cast<size_t>(chunks) * static_cast<size_t>(k);
}

cudaError_t launch_search_topk(const float* logits, long long ld, int n_queries, int N, int k, float scale, float bias,
                               int activation, float2* stats, unsigned long long* keys_a, unsigned long long* keys_b,
                               long long* top_index, float* top_prob, cudaStream_t st) {
  if (n_queries <= 0 || N <= 0 || k <= 0) return cudaSuccess;
  if (k > kChunk / 2 || k > N) return cudaErrorInvalidValue;
  search_stats_kernel<<<n_queries, 1024, 0, st>>>(logits, ld, N, scale, bias, stats);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const long long per_query = static_cast<long long>(search_scratch_keys(N, k));
  int n_in = N;
  bool first = true;
  unsigned long long *in = nullptr, *out = keys_a;
  for (;;) {
    const int chunks = (n_in + kChunk - 1) / kChunk;
    const dim3 grid(chunks, n_queries);
    if (chunks == 1) {
      if (first)
        topk_round_kernel<true, true><<<grid, kSortThreads, 0, st>>>(logits, ld, nullptr, 0, n_in, scale, bias, k, nullptr, 0,
                                                                   stats, activation, top_index, top_prob, k);
      else
        topk_round_kernel<false, true><<<grid, kSortThreads, 0, st>>>(nullptr, 0, in, per_query, n_in, scale, bias, k, nullptr,
                                                                    0, stats, activation, top_index, top_prob, k);
      return cudaGetLastError();
    }
    if (first)
      topk_round_kernel<true, false><<<grid, kSortThreads, 0, st>>>(logits, ld, nullptr, 0, n_in, scale, bias, k, out,
                                                                  per_query, stats, activation, nullptr, nullptr, k);
    else
      topk_round_kernel<false, false><<<grid, kSortThreads, 0, st>>>(nullptr, 0, in, per_query, n_in, scale, bias, k, out,
                                                                   per_query, stats, activation, nullptr, nullptr, k);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    n_in = chunks * k;   // every chunk contributed k keys (padding keys are 0 and sort last)
    first = false;
    in = out;
    out = (out == keys_a) ? keys_b : keys_a;
  }
}

}  // namespace clipb200
