// Convolution-stage kernels for the MobileCLIP2 / FastViT hybrid trunk (NHWC activations, channels contiguous):
//   stem_conv3x3_s2     dense 3x3 stride-2 conv on the raw image (uint8 through the normalisation LUT, or f32 NCHW)
//   dwconv              depthwise KxK (K = 3 | 7), stride 1 | 2, channel multiplier 1 | 2, fused bias (+ GELU)
//   gap                 global average pool over the pixels of each image
//   se_mlp              squeeze-excite gate: sigmoid(W2 relu(W1 s + b1) + b2)
//   scale_act           x * gate[b, c] (+ GELU), fp32 -> fp32 | bf16
//   cast_affine         fp32 -> bf16 copy (input of the 1x1-conv GEMMs)
// These replace the Conv / GlobalAveragePool / Sigmoid / Mul / Erf nodes ONNX Runtime runs for the re-parameterised
// FastViT graph (reference pull_onnx.py:110-116; SURVEY.md Appendix A "C2").  1x1 convolutions and the attention
// blocks of the last stage run on the tcgen05 GEMM / attention kernels.  All of these are HBM/L2-bound.
#include "conv_kernels.cuh"

#include <math.h>
#include <stdlib.h>

namespace clipb200 {

// same Abramowitz-Stegun form as gemm_sm100.cuh::gelu_erf_fast (|error| <= 4.3e-7)
__device__ __forceinline__ float gelu_erf(float x) {
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.23164189f, fabsf(x), 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752f));
  float q = fmaf(t, 0.5307027145f, -0.7265760135f);
  q = fmaf(t, q, 0.7107068705f);
  q = fmaf(t, q, -0.142248368f);
  q = fmaf(t, q, 0.127414796f);
  const float h = q * t * e;
  return x * (x < 0.f ? h : 1.0f - h);
}
__device__ __forceinline__ float ld_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_from_float(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_from_float(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

// ------------------------------------------------------------------------------------------------- stem
// out[b, oy, ox, oc] = gelu(bias[oc] + sum_{c,ky,kx} w[oc][c][ky][kx] * in[b, 2*oy+ky-1, 2*ox+kx-1, c]); pad 1.
// One thread = one output pixel x 8 output channels; the 27 taps are read once per thread.
__global__ void __launch_bounds__(256)
stem_conv3x3_s2_kernel(const uint8_t* __restrict__ img_u8, const float* __restrict__ img_f32, const float* __restrict__ lut,
                       int S, int Cout, const float* __restrict__ w /*[27][Cout]*/, const float* __restrict__ bias,
                       long long total, __nv_bfloat16* __restrict__ out) {
  extern __shared__ float stem_smem[];  // [27*Cout] weights + [Cout] bias + [768] LUT
  float* sw = stem_smem;
  float* sb = sw + 27 * Cout;
  float* sl = sb + Cout;
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sb[i] = bias[i];
  if (img_u8 != nullptr)
    for (int i = threadIdx.x; i < 768; i += blockDim.x) sl[i] = lut[i];
  __syncthreads();
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int cg = Cout >> 3;
  const int g = static_cast<int>(idx % cg);
  const long long pix = idx / cg;
  const int So = S >> 1;
  const int ox = static_cast<int>(pix % So), oy = static_cast<int>((pix / So) % So);
  const long long b = pix / (static_cast<long long>(So) * So);
  float taps[27];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int iy = 2 * oy + ky - 1, ix = 2 * ox + kx - 1;
      const bool ok = iy >= 0 && iy < S && ix >= 0 && ix < S;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v = 0.f;
        if (ok) {
          if (img_u8 != nullptr) v = sl[c * 256 + img_u8[((b * S + iy) * S + ix) * 3 + c]];
          else v = __ldg(img_f32 + ((b * 3 + c) * S + iy) * S + ix);
        }
        taps[c * 9 + ky * 3 + kx] = v;
      }
    }
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = sb[g * 8 + e];
#pragma unroll
  for (int t = 0; t < 27; ++t)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(taps[t], sw[t * Cout + g * 8 + e], acc[e]);
  __nv_bfloat162 h[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(gelu_erf(acc[2 * e]), gelu_erf(acc[2 * e + 1]));
  *reinterpret_cast<uint4*>(out + pix * Cout + g * 8) = *reinterpret_cast<uint4*>(h);
}

// Same convolution, one thread = one output pixel x ALL output channels (Cout % 16 == 0): the 27 taps are gathered once
// per pixel (the kernel above gathers them Cout/8 times), weights come from shared memory as warp-broadcast 16-byte
// loads (27 x 4 LDS.128 per 16 channels x 27 FFMA), and the pixel's Cout bf16 outputs leave as 16-byte stores.
__global__ void __launch_bounds__(128)
stem_conv3x3_s2_pix_kernel(const uint8_t* __restrict__ img_u8, const float* __restrict__ img_f32,
                           const float* __restrict__ lut, int S, int Cout, const float* __restrict__ w /*[27][Cout]*/,
                           const float* __restrict__ bias, long long total_pix, __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) float stem_smem[];  // [27*Cout] weights + [Cout] bias + [768] LUT
  float* sw = stem_smem;
  float* sb = sw + 27 * Cout;
  float* sl = sb + Cout;
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sb[i] = bias[i];
  if (img_u8 != nullptr)
    for (int i = threadIdx.x; i < 768; i += blockDim.x) sl[i] = lut[i];
  __syncthreads();
  const long long pix = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (pix >= total_pix) return;
  const int So = S >> 1;
  const int ox = static_cast<int>(pix % So), oy = static_cast<int>((pix / So) % So);
  const long long b = pix / (static_cast<long long>(So) * So);
  float taps[27];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int iy = 2 * oy + ky - 1, ix = 2 * ox + kx - 1;
      const bool ok = iy >= 0 && iy < S && ix >= 0 && ix < S;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v = 0.f;
        if (ok) {
          if (img_u8 != nullptr) v = sl[c * 256 + __ldg(img_u8 + ((b * S + iy) * S + ix) * 3 + c)];
          else v = __ldg(img_f32 + ((b * 3 + c) * S + iy) * S + ix);
        }
        taps[c * 9 + ky * 3 + kx] = v;
      }
    }
  __nv_bfloat16* orow = out + pix * Cout;
  for (int c0 = 0; c0 < Cout; c0 += 16) {
    float4 a[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) a[q] = *reinterpret_cast<const float4*>(sb + c0 + 4 * q);
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const float v = taps[t];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 w4 = *reinterpret_cast<const float4*>(sw + t * Cout + c0 + 4 * q);
        a[q].x = fmaf(v, w4.x, a[q].x); a[q].y = fmaf(v, w4.y, a[q].y);
        a[q].z = fmaf(v, w4.z, a[q].z); a[q].w = fmaf(v, w4.w, a[q].w);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; q += 2) {
      __nv_bfloat162 h[4];
      h[0] = __floats2bfloat162_rn(gelu_erf(a[q].x), gelu_erf(a[q].y));
      h[1] = __floats2bfloat162_rn(gelu_erf(a[q].z), gelu_erf(a[q].w));
      h[2] = __floats2bfloat162_rn(gelu_erf(a[q + 1].x), gelu_erf(a[q + 1].y));
      h[3] = __floats2bfloat162_rn(gelu_erf(a[q + 1].z), gelu_erf(a[q + 1].w));
      *reinterpret_cast<uint4*>(orow + c0 + 4 * q) = *reinterpret_cast<uint4*>(h);
    }
  }
}

cudaError_t launch_stem_conv3x3_s2(const uint8_t* img_u8, const float* img_f32, const float* lut, int n, int S, int Cout,
                                   const float* w27, const float* bias, __nv_bfloat16* out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if ((Cout & 7) || (S & 1)) return cudaErrorInvalidValue;
  const size_t smem = (27 * Cout + Cout + 768) * sizeof(float);
  if ((Cout & 15) == 0) {
    const long long pixels = static_cast<long long>(n) * (S / 2) * (S / 2);
    stem_conv3x3_s2_pix_kernel<<<static_cast<unsigned>((pixels + 127) / 128), 128, smem, st>>>(img_u8, img_f32, lut, S, Cout,
                                                                                               w27, bias, pixels, out);
    return cudaGetLastError();
  }
  const long long total = static_cast<long long>(n) * (S / 2) * (S / 2) * (Cout / 8);
  stem_conv3x3_s2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, smem, st>>>(img_u8, img_f32, lut, S, Cout, w27,
                                                                                        bias, total, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------- depthwise
// in [B,H,W,Cin], out [B,Ho,Wo,Cin*MULT]; weight rearranged to [K*K][Cout] (channel contiguous).
// Block = one image x TILE x TILE output pixels x 32 output channels; the input halo tile is staged in shared memory
// (coalesced 128-byte channel rows), each thread then slides along one output row keeping its taps in registers.
template <int K, int STRIDE, int MULT, typename Tin, typename Tout, bool GELU>
__global__ void __launch_bounds__(256)
dwconv_kernel(const Tin* __restrict__ in, int H, int W, int Cin, const float* __restrict__ w, const float* __restrict__ bias,
              Tout* __restrict__ out, int Ho, int Wo, int tiles_x) {
  constexpr int TILE = 8;
  constexpr int IT = (TILE - 1) * STRIDE + K;  // input tile edge
  constexpr int CI = 32 / MULT;                // input channels per block
  extern __shared__ float dw_smem[];           // [IT][IT][CI]
  const int Cout = Cin * MULT;
  const int c0 = blockIdx.y * 32;              // first output channel of this block
  const int ci0 = c0 / MULT;
  const int b = blockIdx.z;
  const int ty0 = (blockIdx.x / tiles_x) * TILE, tx0 = (blockIdx.x % tiles_x) * TILE;
  const int iy0 = ty0 * STRIDE - K / 2, ix0 = tx0 * STRIDE - K / 2;
  const Tin* inb = in + static_cast<long long>(b) * H * W * Cin;
  for (int i = threadIdx.x; i < IT * IT * CI; i += 256) {
    const int c = i % CI, p = i / CI;
    const int ix = ix0 + p % IT, iy = iy0 + p / IT;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W && ci0 + c < Cin) v = ld_as_float(inb + (static_cast<long long>(iy) * W + ix) * Cin + ci0 + c);
    dw_smem[i] = v;
  }
  __syncthreads();
  const int lane_c = threadIdx.x & 31;  // output channel within the block
  const int row = threadIdx.x >> 5;     // output row within the tile (8 rows, 8 warps)
  const int oc = c0 + lane_c;
  if (oc >= Cout) return;
  const int cl = lane_c / MULT;         // input channel within the staged tile
  float wk[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) wk[t] = __ldg(w + t * Cout + oc);
  const float bv = __ldg(bias + oc);
  const int oy = ty0 + row;
  if (oy >= Ho) return;
  // sliding window in registers: one staged input row (IT values) serves all TILE outputs of this thread's output
  // row, so shared-memory loads per output drop from K*K to K*IT/TILE
  float acc[TILE];
#pragma unroll
  for (int x = 0; x < TILE; ++x) acc[x] = bv;
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
    float rv[IT];
#pragma unroll
    for (int i = 0; i < IT; ++i) rv[i] = dw_smem[((row * STRIDE + ky) * IT + i) * CI + cl];
#pragma unroll
    for (int x = 0; x < TILE; ++x)
#pragma unroll
      for (int kx = 0; kx < K; ++kx) acc[x] = fmaf(rv[x * STRIDE + kx], wk[ky * K + kx], acc[x]);
  }
#pragma unroll
  for (int x = 0; x < TILE; ++x) {
    const int ox = tx0 + x;
    if (ox < Wo) {
      const float v = GELU ? gelu_erf(acc[x]) : acc[x];
      st_from_float(out + ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * Cout + oc, v);
    }
  }
}

template <int K, int STRIDE, int MULT, typename Tin, typename Tout, bool GELU>
static cudaError_t dw_launch(const Tin* in, int n, int H, int W, int Cin, const float* w, const float* bias, Tout* out,
                             cudaStream_t st) {
  const int Ho = (H + 2 * (K / 2) - K) / STRIDE + 1, Wo = (W + 2 * (K / 2) - K) / STRIDE + 1;
  const int tiles_x = (Wo + 7) / 8, tiles_y = (Ho + 7) / 8;
  constexpr int IT = 7 * STRIDE + K;
  const size_t smem = static_cast<size_t>(IT) * IT * (32 / MULT) * sizeof(float);
  dim3 grid(tiles_x * tiles_y, (Cin * MULT + 31) / 32, n);
  dwconv_kernel<K, STRIDE, MULT, Tin, Tout, GELU><<<grid, 256, smem, st>>>(in, H, W, Cin, w, bias, out, Ho, Wo, tiles_x);
  return cudaGetLastError();
}

cudaError_t launch_dwconv(const void* in, bool in_bf16, int n, int H, int W, int Cin, int K, int stride, int mult,
                          const float* w, const float* bias, bool gelu, void* out, bool out_bf16, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  static const bool legacy = getenv("CLIPB200_DWCONV_LEGACY") != nullptr && atoi(getenv("CLIPB200_DWCONV_LEGACY")) != 0;
  if (!legacy && dwconv_tma_supported(in_bf16, Cin, K, stride, mult, gelu))
    return launch_dwconv_tma(static_cast<const float*>(in), n, H, W, Cin, K, w, bias, out, out_bf16, st);
  if (!legacy) {
    cudaError_t e = cudaSuccess;
    if (launch_dwconv_tma_gen(in, in_bf16, n, H, W, Cin, K, stride, mult, w, bias, gelu, out, out_bf16, st, &e)) return e;
  }
#define CLIPB200_DW(K_, S_, M_, TI, TO, G_)                                                                     \
  if (K == K_ && stride == S_ && mult == M_ && in_bf16 == std::is_same<TI, __nv_bfloat16>::value &&             \
      out_bf16 == std::is_same<TO, __nv_bfloat16>::value && gelu == G_)                                         \
    return dw_launch<K_, S_, M_, TI, TO, G_>(static_cast<const TI*>(in), n, H, W, Cin, w, bias, static_cast<TO*>(out), st);
  // the combinations the FastViT trunk uses
  CLIPB200_DW(3, 2, 1, __nv_bfloat16, __nv_bfloat16, true)   // stem.1
  CLIPB200_DW(3, 1, 1, float, float, false)                  // RepMixer token mixer
  CLIPB200_DW(7, 1, 1, float, __nv_bfloat16, false)          // ConvMlp depthwise (feeds the fc1 GEMM)
  CLIPB200_DW(7, 1, 1, float, float, false)                  // RepCPE positional encoding
  CLIPB200_DW(7, 2, 2, float, float, false)                  // downsample large-kernel conv (SE follows)
  CLIPB200_DW(7, 2, 2, float, __nv_bfloat16, true)           // downsample large-kernel conv + GELU (no SE)
  CLIPB200_DW(3, 1, 2, float, float, false)                  // final_conv (SE follows)
#undef CLIPB200_DW
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------------- GAP / SE
// mean over the P pixels of each image: grid (C/32 column groups, B); 8 warps stride over pixels, lanes over channels.
__global__ void __launch_bounds__(256)
gap_kernel(const float* __restrict__ x, int P, int C, float* __restrict__ out) {
  __shared__ float part[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane, b = blockIdx.y;
  float acc = 0.f;
  if (c < C)
    for (int p = warp; p < P; p += 8) acc += __ldg(x + (static_cast<long long>(b) * P + p) * C + c);
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i][lane];
    out[static_cast<long long>(b) * C + c] = s / static_cast<float>(P);
  }
}
cudaError_t launch_gap(const float* x, int n, int P, int C, float* out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  gap_kernel<<<dim3((C + 31) / 32, n), 256, 0, st>>>(x, P, C, out);
  return cudaGetLastError();
}

// gate[b, c] = sigmoid(b2[c] + sum_r w2[c][r] * relu(b1[r] + sum_k w1[r][k] * s[b, k])); one block per image
__global__ void __launch_bounds__(256)
se_mlp_kernel(const float* __restrict__ s, int C, int R, const float* __restrict__ w1, const float* __restrict__ b1,
              const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ gate) {
  extern __shared__ float se_smem[];  // [C] squeezed + [R] hidden
  float* ss = se_smem;
  float* sh = ss + C;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < C; i += 256) ss[i] = s[static_cast<long long>(b) * C + i];
  __syncthreads();
  for (int r = warp; r < R; r += 8) {
    float acc = 0.f;
    for (int k = lane; k < C; k += 32) acc = fmaf(__ldg(w1 + static_cast<long long>(r) * C + k), ss[k], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) sh[r] = fmaxf(acc + __ldg(b1 + r), 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float acc = __ldg(b2 + c);
    for (int r = 0; r < R; ++r) acc = fmaf(__ldg(w2 + static_cast<long long>(c) * R + r), sh[r], acc);
    gate[static_cast<long long>(b) * C + c] = 1.0f / (1.0f + expf(-acc));
  }
}
cudaError_t launch_se_mlp(const float* s, int n, int C, int R, const float* w1, const float* b1, const float* w2,
                          const float* b2, float* gate, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  se_mlp_kernel<<<n, 256, (C + R) * sizeof(float), st>>>(s, C, R, w1, b1, w2, b2, gate);
  return cudaGetLastError();
}

// out[b, p, c] = act(x[b, p, c] * gate[b, c]); gate may be null (plain cast / activation)
template <typename Tout>
__global__ void __launch_bounds__(256)
scale_act_kernel(const float* __restrict__ x, const float* __restrict__ gate, long long total4, int P, int C, int gelu,
                 Tout* __restrict__ out) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total4) return;
  const long long e = i * 4;
  const int c = static_cast<int>(e % C);
  const long long b = e / (static_cast<long long>(P) * C);
  float4 v = *reinterpret_cast<const float4*>(x + e);
  if (gate != nullptr) {
    const float4 g = *reinterpret_cast<const float4*>(gate + b * C + c);
    v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
  }
  if (gelu) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
  st_from_float(out + e, v.x); st_from_float(out + e + 1, v.y); st_from_float(out + e + 2, v.z); st_from_float(out + e + 3, v.w);
}
cudaError_t launch_scale_act(const float* x, const float* gate, int n, int P, int C, bool gelu, void* out, bool out_bf16,
                             cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (C & 3) return cudaErrorInvalidValue;
  const long long total4 = static_cast<long long>(n) * P * C / 4;
  const unsigned blocks = static_cast<unsigned>((total4 + 255) / 256);
  if (out_bf16) scale_act_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(x, gate, total4, P, C, gelu ? 1 : 0, static_cast<__nv_bfloat16*>(out));
  else scale_act_kernel<float><<<blocks, 256, 0, st>>>(x, gate, total4, P, C, gelu ? 1 : 0, static_cast<float*>(out));
  return cudaGetLastError();
}

}  // namespace clipb200
