// See kernels.cuh for the inventory.  sm_100a only.
#include "kernels.cuh"

#include <math.h>

namespace clipb200 {

// =================================================================================================
// small helpers
// =================================================================================================
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// =================================================================================================
// preprocess: u8 HWC -> normalised bf16 patch matrix  (vision.rs:235-259 fused with the conv's im2col)
//   patches[(n*G*G + py*G + px), k],  k = c*P*P + iy*P + ix  (== conv weight [D,3,P,P] flattened)
//   value = LUT[c][byte]  with LUT[c][v] = ((float)v / 255.0f - mean[c]) / std[c] built on the host in fp32
//   (bit-identical to the reference expression), then rounded once to bf16.
// One block = one band of P image rows (contiguous P*S*3 bytes): coalesced word loads into shared memory,
// 16-byte coalesced stores of the patch rows.
// =================================================================================================
template <int P>
__global__ void __launch_bounds__(256)
preprocess_patches_u8_kernel(const uint8_t* __restrict__ img, long long total_bytes, int S, int G, int Pr, int K,
                             int Kp, const float* __restrict__ lut, __nv_bfloat16* __restrict__ patches) {
  extern __shared__ __align__(16) uint8_t pp_smem[];
  const int PP = (P > 0) ? P : Pr;
  float* slut = reinterpret_cast<float*>(pp_smem);
  uint8_t* band = pp_smem + 3 * 256 * sizeof(float);
  const int py = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
  const long long start = (static_cast<long long>(n) * S + static_cast<long long>(py) * PP) * S * 3;
  const int band_bytes = PP * S * 3;
  for (int i = tid; i < 768; i += blockDim.x) slut[i] = lut[i];
  const long long a0 = start & ~3ll;
  const int shift = static_cast<int>(start - a0);
  const int nwords = (shift + band_bytes + 3) >> 2;
  for (int w = tid; w < nwords; w += blockDim.x) {
    const long long off = a0 + 4ll * w;
    uint32_t v;
    if (off + 4 <= total_bytes) {
      v = __ldg(reinterpret_cast<const uint32_t*>(img + off));
    } else {
      v = 0;
      for (int j = 0; j < 4; ++j)
        if (off + j < total_bytes) v |= static_cast<uint32_t>(img[off + j]) << (8 * j);
    }
    reinterpret_cast<uint32_t*>(band)[w] = v;
  }
  __syncthreads();
  const uint8_t* b = band + shift;
  const int chunks = Kp >> 3;
  const long long row0 = (static_cast<long long>(n) * G + py) * G;
  for (int idx = tid; idx < G * chunks; idx += blockDim.x) {
    const int px = idx / chunks, ch = idx - px * chunks;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = ch * 8 + e;
      if (k < K) {
        const int c = k / (PP * PP);
        const int rem = k - c * PP * PP;
        const int iy = rem / PP, ix = rem - iy * PP;
        const uint8_t byte = b[(iy * S + px * PP + ix) * 3 + c];
        v[e] = slut[c * 256 + byte];
      } else {
        v[e] = 0.f;
      }
    }
    uint4 pk;
    pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
    pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(patches + (row0 + px) * Kp + ch * 8) = pk;
  }
}

cudaError_t launch_preprocess_patches_u8(const uint8_t* img, int n, int S, int P, int Kp, const float* lut,
                                         __nv_bfloat16* patches, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (S % P != 0 || (Kp & 7)) return cudaErrorInvalidValue;
  const int G = S / P, K = 3 * P * P;
  const size_t smem = 3 * 256 * sizeof(float) + static_cast<size_t>(P) * S * 3 + 8;
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  const long long total = static_cast<long long>(n) * S * S * 3;
  dim3 grid(G, n);
  switch (P) {
    case 14: preprocess_patches_u8_kernel<14><<<grid, 256, smem, st>>>(img, total, S, G, P, K, Kp, lut, patches); break;
    case 16: preprocess_patches_u8_kernel<16><<<grid, 256, smem, st>>>(img, total, S, G, P, K, Kp, lut, patches); break;
    case 32: preprocess_patches_u8_kernel<32><<<grid, 256, smem, st>>>(img, total, S, G, P, K, Kp, lut, patches); break;
    default: preprocess_patches_u8_kernel<0><<<grid, 256, smem, st>>>(img, total, S, G, P, K, Kp, lut, patches); break;
  }
  return cudaGetLastError();
}

// u8 HWC -> f32 CHW through the same LUT: the reference's public `preprocess` output (vision.rs:120-140), bit-exact.
__global__ void __launch_bounds__(256)
normalize_nchw_f32_kernel(const uint8_t* __restrict__ img, int S, long long total, const float* __restrict__ lut,
                          float* __restrict__ out) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const long long plane = static_cast<long long>(S) * S;
  const long long n = idx / (3 * plane);
  const long long rem = idx - n * 3 * plane;
  const int c = static_cast<int>(rem / plane);
  const long long i = rem - c * plane;
  out[idx] = __ldg(lut + c * 256 + img[(n * plane + i) * 3 + c]);
}
cudaError_t launch_normalize_nchw_f32(const uint8_t* img, int n, int S, const float* lut, float* out, cudaStream_t st) {
  const long long total = static_cast<long long>(n) * 3 * S * S;
  if (total <= 0) return cudaSuccess;
  normalize_nchw_f32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(img, S, total, lut, out);
  return cudaGetLastError();
}

// f32 NCHW pixel_values (what the reference hands to ORT, vision.rs:105) -> bf16 patch matrix
__global__ void __launch_bounds__(256)
im2col_f32_kernel(const float* __restrict__ x, int S, int G, int P, int K, int Kp, long long total_chunks,
                  __nv_bfloat16* __restrict__ patches) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total_chunks) return;
  const int chunks = Kp >> 3;
  const long long row = idx / chunks;
  const int ch = static_cast<int>(idx - row * chunks);
  const int px = static_cast<int>(row % G);
  const int py = static_cast<int>((row / G) % G);
  const long long n = row / (static_cast<long long>(G) * G);
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = ch * 8 + e;
    if (k < K) {
      const int c = k / (P * P);
      const int rem = k - c * P * P;
      const int iy = rem / P, ix = rem - iy * P;
      v[e] = __ldg(x + ((n * 3 + c) * S + (py * P + iy)) * S + px * P + ix);
    } else {
      v[e] = 0.f;
    }
  }
  uint4 pk;
  pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
  pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(patches + row * Kp + ch * 8) = pk;
}

cudaError_t launch_im2col_f32(const float* nchw, int n, int S, int P, int Kp, __nv_bfloat16* patches,
                              cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (S % P != 0 || (Kp & 7)) return cudaErrorInvalidValue;
  const int G = S / P;
  const long long total = static_cast<long long>(n) * G * G * (Kp >> 3);
  im2col_f32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(nchw, S, G, P, 3 * P * P, Kp, total,
                                                                                 patches);
  return cudaGetLastError();
}

// =================================================================================================
// LayerNorm: one warp per row, values held in registers, two-pass (mean, then centred variance) in fp32.
// =================================================================================================
constexpr int LN_MAX_VEC = 16;  // float4 per lane -> D <= 2048

// LN_MAX_VEC is a template parameter so that a 1152-wide row keeps 9 float4 per lane in registers, not 16:
// fewer registers -> more resident warps -> more loads in flight (this kernel is purely HBM-bound).
template <int LN_MAX_VEC>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const int* __restrict__ row_map, int rows, int D,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                 __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32, int descending) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  // descending: the blocks scheduled first take the LAST rows — the ones the preceding reduce-add GEMM wrote last and
  // that are still in L2 — and the rows normalised last are the first ones, which the following GEMM reads first
  if (descending) row = rows - 1 - row;
  const long long src = row_map != nullptr ? row_map[row] : row;
  const float4* xr = reinterpret_cast<const float4*>(x + src * D);
  const int nvec = D >> 2;
  float4 v[LN_MAX_VEC];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAX_VEC; ++j) {
    const int i = lane + 32 * j;
    if (i < nvec) {
      v[j] = xr[i];
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
  }
  const float mean = warp_sum(s) / static_cast<float>(D);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAX_VEC; ++j) {
    const int i = lane + 32 * j;
    if (i < nvec) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(D) + eps);
#pragma unroll
  for (int j = 0; j < LN_MAX_VEC; ++j) {
    const int i = lane + 32 * j;
    if (i < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(beta) + i);
      float4 y;
      y.x = (v[j].x - mean) * rstd * g.x + bb.x;
      y.y = (v[j].y - mean) * rstd * g.y + bb.y;
      y.z = (v[j].z - mean) * rstd * g.z + bb.z;
      y.w = (v[j].w - mean) * rstd * g.w + bb.w;
      if (out_bf16 != nullptr) {
        uint2 pk;
        pk.x = pack_bf16x2(y.x, y.y);
        pk.y = pack_bf16x2(y.z, y.w);
        reinterpret_cast<uint2*>(out_bf16 + static_cast<long long>(row) * D)[i] = pk;
      } else {
        reinterpret_cast<float4*>(out_f32 + static_cast<long long>(row) * D)[i] = y;
      }
    }
  }
}

cudaError_t launch_layernorm(const float* x, const int* row_map, int rows, int D, const float* gamma,
                             const float* beta, float eps, __nv_bfloat16* out_bf16, float* out_f32, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  if ((D & 3) || D > LN_MAX_VEC * 128) return cudaErrorInvalidValue;
  const int nv = (D / 4 + 31) / 32;
  const dim3 grid((rows + 7) / 8);
  // CLIPB200_LN_DESCENDING=0 restores ascending order (A/B: DFN5B text LayerNorm 41.2 -> 39.0 ms per step, SO400M 43.5 -> 42.9;
  // profiles/r02am_ln_order.log)
  static const int desc = [] { const char* v = getenv("CLIPB200_LN_DESCENDING"); return v == nullptr || atoi(v) != 0 ? 1 : 0; }();
#define CLIPB200_LN_CASE(NV_)                                                                                     \
  if (nv <= NV_) {                                                                                                \
    layernorm_kernel<NV_><<<grid, 256, 0, st>>>(x, row_map, rows, D, gamma, beta, eps, out_bf16, out_f32, desc);  \
    return cudaGetLastError();                                                                                    \
  }
  CLIPB200_LN_CASE(2)
  CLIPB200_LN_CASE(4)
  CLIPB200_LN_CASE(6)
  CLIPB200_LN_CASE(8)
  CLIPB200_LN_CASE(9)
  CLIPB200_LN_CASE(10)
  CLIPB200_LN_CASE(12)
  CLIPB200_LN_CASE(16)
#undef CLIPB200_LN_CASE
  return cudaErrorInvalidValue;
}

// =================================================================================================
// Flash attention (forward, no dropout), bf16 in / bf16 out, fp32 softmax and accumulation.
// grid = (ceil(T/64), H, B), 4 warps x 16 query rows, K/V streamed in 64-key blocks with cp.async double buffering,
// QK^T and PV on mma.sync.m16n8k16 (legacy tensor path; the tcgen05 version is a later round's work).
// =================================================================================================
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int HD, int HDP>
struct FaCfg {
  static constexpr int LDS = HDP + 8;
  static constexpr int SMEM_BYTES = 5 * 64 * LDS * 2;
};

template <int HD, int HDP, bool CAUSAL>
__global__ void __launch_bounds__(128)
flash_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int T, int H,
                       float scale_log2e) {
  constexpr int LDS = FaCfg<HD, HDP>::LDS;
  constexpr int CH = HD / 8;
  extern __shared__ __align__(16) uint8_t fa_smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(fa_smem);
  __nv_bfloat16* sK = sQ + 64 * LDS;
  __nv_bfloat16* sV = sK + 2 * 64 * LDS;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long D3 = 3ll * H * HD;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * T * D3;
  const __nv_bfloat16* gQ = base + h * HD;
  const __nv_bfloat16* gK = base + H * HD + h * HD;
  const __nv_bfloat16* gV = base + 2 * H * HD + h * HD;
  const int q0 = qt * 64;

  // zero the padding columns [HD, LDS) of all five tiles once (cp.async never touches them)
  constexpr int PADC = LDS - HD;
  for (int i = tid; i < 5 * 64 * PADC; i += 128) {
    const int r = i / PADC, c = HD + i % PADC;
    sQ[r * LDS + c] = __float2bfloat16(0.f);
  }
  auto load_tile = [&](__nv_bfloat16* s, const __nv_bfloat16* g, int r0) {
    for (int i = tid; i < 64 * CH; i += 128) {
      const int r = i / CH, c = i - r * CH;
      __nv_bfloat16* dst = s + r * LDS + c * 8;
      if (r0 + r < T) cp_async16(dst, g + (r0 + r) * D3 + c * 8);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
  };
  const int kv_blocks = (T + 63) / 64;
  const int n_kv = CAUSAL ? (qt + 1 < kv_blocks ? qt + 1 : kv_blocks) : kv_blocks;
  load_tile(sQ, gQ, q0);
  load_tile(sK, gK, 0);
  load_tile(sV, gV, 0);
  cp_async_commit();

  uint32_t qf[HDP / 16][4];
  float o[HDP / 8][4];
#pragma unroll
  for (int i = 0; i < HDP / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int mi = lane >> 3, lr = lane & 7;
  const int qrow0 = q0 + warp * 16 + (lane >> 2);  // this thread's rows: qrow0 and qrow0 + 8

  for (int kb = 0; kb < n_kv; ++kb) {
    if (kb + 1 < n_kv) {
      load_tile(sK + ((kb + 1) & 1) * 64 * LDS, gK, (kb + 1) * 64);
      load_tile(sV + ((kb + 1) & 1) * 64 * LDS, gV, (kb + 1) * 64);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kb == 0) {
#pragma unroll
      for (int ks = 0; ks < HDP / 16; ++ks)
        ldmatrix_x4(qf[ks], sQ + (warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8);
    }
    const __nv_bfloat16* k_s = sK + (kb & 1) * 64 * LDS;
    const __nv_bfloat16* v_s = sV + (kb & 1) * 64 * LDS;
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < HDP / 16; ++ks) {
#pragma unroll
      for (int nb2 = 0; nb2 < 4; ++nb2) {
        uint32_t bf[4];
        ldmatrix_x4(bf, k_s + (nb2 * 16 + (mi >> 1) * 8 + lr) * LDS + ks * 16 + (mi & 1) * 8);
        mma_bf16_16816(s[2 * nb2], qf[ks], bf[0], bf[1]);
        mma_bf16_16816(s[2 * nb2 + 1], qf[ks], bf[2], bf[3]);
      }
    }
    // scale, mask, online softmax
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kb * 64 + nb * 8 + (lane & 3) * 2 + (e & 1);
        const int qrow = qrow0 + (e >> 1) * 8;
        float val = s[nb][e] * scale_log2e;
        if (key >= T || (CAUSAL && key > qrow)) val = -INFINITY;
        s[nb][e] = val;
        if (e < 2) mx0 = fmaxf(mx0, val); else mx1 = fmaxf(mx1, val);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float base0 = (mn0 == -INFINITY) ? 0.f : mn0, base1 = (mn1 == -INFINITY) ? 0.f : mn1;
    const float corr0 = fast_exp2(m0 - base0), corr1 = fast_exp2(m1 - base1);
    m0 = mn0; m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      s[nb][0] = fast_exp2(s[nb][0] - base0); s[nb][1] = fast_exp2(s[nb][1] - base0);
      s[nb][2] = fast_exp2(s[nb][2] - base1); s[nb][3] = fast_exp2(s[nb][3] - base1);
      rs0 += s[nb][0] + s[nb][1];
      rs1 += s[nb][2] + s[nb][3];
    }
    l0 = l0 * corr0 + rs0;
    l1 = l1 * corr1 + rs1;
#pragma unroll
    for (int i = 0; i < HDP / 8; ++i) { o[i][0] *= corr0; o[i][1] *= corr0; o[i][2] *= corr1; o[i][3] *= corr1; }
    // O += P V
#pragma unroll
    for (int ks2 = 0; ks2 < 4; ++ks2) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * ks2][0], s[2 * ks2][1]);
      pa[1] = pack_bf16x2(s[2 * ks2][2], s[2 * ks2][3]);
      pa[2] = pack_bf16x2(s[2 * ks2 + 1][0], s[2 * ks2 + 1][1]);
      pa[3] = pack_bf16x2(s[2 * ks2 + 1][2], s[2 * ks2 + 1][3]);
#pragma unroll
      for (int nb2 = 0; nb2 < HDP / 16; ++nb2) {
        uint32_t bf[4];
        ldmatrix_x4_trans(bf, v_s + (ks2 * 16 + (mi & 1) * 8 + lr) * LDS + nb2 * 16 + (mi >> 1) * 8);
        mma_bf16_16816(o[2 * nb2], pa, bf[0], bf[1]);
        mma_bf16_16816(o[2 * nb2 + 1], pa, bf[2], bf[3]);
      }
    }
    __syncthreads();  // the buffer read here is overwritten by the prefetch of iteration kb+1
  }
  // finalise: quad-reduce the row sums, normalise, stage through this warp's own 16 rows of sQ, coalesced store
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
  __nv_bfloat16* stage = sQ + warp * 16 * LDS;
#pragma unroll
  for (int nb = 0; nb < HDP / 8; ++nb) {
    const int col = nb * 8 + (lane & 3) * 2;
    *reinterpret_cast<uint32_t*>(stage + (lane >> 2) * LDS + col) = pack_bf16x2(o[nb][0] * inv0, o[nb][1] * inv0);
    *reinterpret_cast<uint32_t*>(stage + ((lane >> 2) + 8) * LDS + col) = pack_bf16x2(o[nb][2] * inv1, o[nb][3] * inv1);
  }
  __syncwarp();
  const long long DO = static_cast<long long>(H) * HD;
  for (int i = lane; i < 16 * CH; i += 32) {
    const int r = i / CH, c = i - r * CH;
    const int q = q0 + warp * 16 + r;
    if (q < T)
      *reinterpret_cast<uint4*>(out + (static_cast<long long>(b) * T + q) * DO + h * HD + c * 8) =
          *reinterpret_cast<const uint4*>(stage + r * LDS + c * 8);
  }
}

template <int HD, int HDP>
static cudaError_t fa_launch(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int H, bool causal,
                             cudaStream_t st) {
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  dim3 grid((T + 63) / 64, H, B);
  constexpr int smem = FaCfg<HD, HDP>::SMEM_BYTES;
  if (causal) flash_attention_kernel<HD, HDP, true><<<grid, 128, smem, st>>>(qkv, out, T, H, scale_log2e);
  else flash_attention_kernel<HD, HDP, false><<<grid, 128, smem, st>>>(qkv, out, T, H, scale_log2e);
  return cudaGetLastError();
}
template <int HD, int HDP>
static cudaError_t fa_configure() {
  constexpr int smem = FaCfg<HD, HDP>::SMEM_BYTES;
  cudaError_t e = cudaFuncSetAttribute(flash_attention_kernel<HD, HDP, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(flash_attention_kernel<HD, HDP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              smem);
}

cudaError_t flash_attention_configure_device() {
  cudaError_t e;
  // The shipped library only instantiates the head dims the tcgen05 kernel (attn_sm100.cuh) does not serve: 32 (FastViT
  // attention stages) and 128.  tests/native/attn_test.cu builds this file with CLIPB200_FA_ALL_HEAD_DIMS to have an
  // independent mma.sync implementation to compare the tcgen05 kernel with.
  if ((e = fa_configure<32, 32>()) != cudaSuccess) return e;
  if ((e = fa_configure<128, 128>()) != cudaSuccess) return e;
#ifdef CLIPB200_FA_ALL_HEAD_DIMS
  if ((e = fa_configure<64, 64>()) != cudaSuccess) return e;
  if ((e = fa_configure<72, 80>()) != cudaSuccess) return e;
  if ((e = fa_configure<80, 80>()) != cudaSuccess) return e;
  if ((e = fa_configure<96, 96>()) != cudaSuccess) return e;
#endif
  return cudaSuccess;
}

cudaError_t launch_flash_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int H, int hd,
                                   bool causal, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  switch (hd) {
    case 32: return fa_launch<32, 32>(qkv, out, B, T, H, causal, st);
    case 128: return fa_launch<128, 128>(qkv, out, B, T, H, causal, st);
#ifdef CLIPB200_FA_ALL_HEAD_DIMS
    case 64: return fa_launch<64, 64>(qkv, out, B, T, H, causal, st);
    case 72: return fa_launch<72, 80>(qkv, out, B, T, H, causal, st);
    case 80: return fa_launch<80, 80>(qkv, out, B, T, H, causal, st);
    case 96: return fa_launch<96, 96>(qkv, out, B, T, H, causal, st);
#endif
    default: return cudaErrorInvalidValue;
  }
}

// =================================================================================================
// SigLIP attention pooling (timm AttentionPoolLatent, latent_len 1): one query per head, T keys.
// grid = (H, B), 128 threads.  HBM-bound: reads the [T, 2*hd] slice of kv once.
// =================================================================================================
__global__ void __launch_bounds__(128)
map_pool_attention_kernel(const __nv_bfloat16* __restrict__ kv, const float* __restrict__ q,
                          __nv_bfloat16* __restrict__ out, int T, int H, int hd) {
  extern __shared__ __align__(16) float mp_smem[];
  float* sp = mp_smem;            // [T] scores / probabilities
  float* sq = sp + ((T + 3) & ~3);  // [hd]
  float* part = sq + 128;         // [G][hd] partial sums (<= 1024 floats)
  __shared__ float red[8];
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long ld = 2ll * H * hd;
  const __nv_bfloat16* kbase = kv + static_cast<long long>(b) * T * ld + h * hd;
  const __nv_bfloat16* vbase = kbase + H * hd;
  const int CH = hd >> 3;
  for (int i = tid; i < hd; i += 128) sq[i] = q[h * hd + i];
  __syncthreads();
  float lmax = -INFINITY;
  for (int t = tid; t < T; t += 128) {
    const uint4* kr = reinterpret_cast<const uint4*>(kbase + t * ld);
    float acc = 0.f;
    for (int c = 0; c < CH; ++c) {
      const uint4 u = __ldg(kr + c);
      const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(p2[e]);
        acc += f.x * sq[c * 8 + 2 * e] + f.y * sq[c * 8 + 2 * e + 1];
      }
    }
    sp[t] = acc;
    lmax = fmaxf(lmax, acc);
  }
  lmax = warp_max(lmax);
  if (lane == 0) red[warp] = lmax;
  __syncthreads();
  const float gmax = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  float lsum = 0.f;
  for (int t = tid; t < T; t += 128) {
    const float e = __expf(sp[t] - gmax);
    sp[t] = e;
    lsum += e;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) red[4 + warp] = lsum;
  __syncthreads();
  const float inv = 1.f / (red[4] + red[5] + red[6] + red[7]);
  const int G = 128 / CH;
  const int g = tid / CH, c = tid - g * CH;
  if (g < G) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int t = g; t < T; t += G) {
      const float p = sp[t];
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(vbase + t * ld) + c);
      const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(p2[e]);
        acc[2 * e] += p * f.x;
        acc[2 * e + 1] += p * f.y;
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) part[g * hd + c * 8 + e] = acc[e];
  }
  __syncthreads();
  for (int d = tid; d < hd; d += 128) {
    float a = 0.f;
    for (int gg = 0; gg < G; ++gg) a += part[gg * hd + d];
    out[static_cast<long long>(b) * H * hd + h * hd + d] = __float2bfloat16(a * inv);
  }
}

cudaError_t launch_map_pool_attention(const __nv_bfloat16* kv, const float* q, __nv_bfloat16* out, int B, int T, int H,
                                      int hd, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  if ((hd & 7) || hd > 128) return cudaErrorInvalidValue;
  const size_t smem = (static_cast<size_t>((T + 3) & ~3) + 128 + 1024) * sizeof(float);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  map_pool_attention_kernel<<<dim3(H, B), 128, smem, st>>>(kv, q, out, T, H, hd);
  return cudaGetLastError();
}

// =================================================================================================
// text-side small kernels
// =================================================================================================
__global__ void __launch_bounds__(128)
embed_tokens_kernel(const int64_t* __restrict__ ids, int ctx, int D, int vocab, const float* __restrict__ tok,
                    const float* __restrict__ pos, float* __restrict__ x, int* __restrict__ err_flag) {
  const long long row = blockIdx.x;
  long long id = ids[row];
  if (id < 0 || id >= vocab) {  // ORT's Gather would fail the run; flag it and clamp so we never read out of bounds
    if (threadIdx.x == 0) atomicExch(err_flag, 1);
    id = 0;
  }
  const float4* tr = reinterpret_cast<const float4*>(tok + id * D);
  const float4* pr = reinterpret_cast<const float4*>(pos + (row % ctx) * D);
  float4* xr = reinterpret_cast<float4*>(x + row * D);
  for (int i = threadIdx.x; i < (D >> 2); i += blockDim.x) {
    const float4 a = __ldg(tr + i), p = __ldg(pr + i);
    xr[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
  }
}
cudaError_t launch_embed_tokens(const int64_t* ids, int rows, int ctx, int D, int vocab, const float* tok_emb,
                                const float* pos_emb, float* x, int* err_flag, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  if (D & 3) return cudaErrorInvalidValue;
  embed_tokens_kernel<<<rows, 128, 0, st>>>(ids, ctx, D, vocab, tok_emb, pos_emb, x, err_flag);
  return cudaGetLastError();
}

__global__ void text_pool_rows_kernel(const int64_t* __restrict__ ids, int B, int ctx, int argmax, int* row_map) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int best = ctx - 1;
  if (argmax) {
    long long bv = ids[static_cast<long long>(b) * ctx];
    best = 0;
    for (int t = 1; t < ctx; ++t) {
      const long long v = ids[static_cast<long long>(b) * ctx + t];
      if (v > bv) { bv = v; best = t; }  // first maximum, like torch.argmax / ONNX ArgMax(select_last_index=0)
    }
  }
  row_map[b] = b * ctx + best;
}
cudaError_t launch_text_pool_rows(const int64_t* ids, int B, int ctx, bool argmax, int* row_map, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  text_pool_rows_kernel<<<(B + 127) / 128, 128, 0, st>>>(ids, B, ctx, argmax ? 1 : 0, row_map);
  return cudaGetLastError();
}

__global__ void affine_rows_kernel(int B, int stride, int offset, int* row_map) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) row_map[b] = b * stride + offset;
}
cudaError_t launch_affine_rows(int B, int stride, int offset, int* row_map, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  affine_rows_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, stride, offset, row_map);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(128)
write_cls_rows_kernel(float* __restrict__ x, int T, int D, const float* __restrict__ cls_row) {
  float4* xr = reinterpret_cast<float4*>(x + static_cast<long long>(blockIdx.x) * T * D);
  const float4* c = reinterpret_cast<const float4*>(cls_row);
  for (int i = threadIdx.x; i < (D >> 2); i += blockDim.x) xr[i] = __ldg(c + i);
}
cudaError_t launch_write_cls_rows(float* x, int B, int T, int D, const float* cls_row, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  write_cls_rows_kernel<<<B, 128, 0, st>>>(x, T, D, cls_row);
  return cudaGetLastError();
}

// =================================================================================================
// L2 normalise (F.normalize: x / max(||x||_2, 1e-12)), one warp per row
// =================================================================================================
__global__ void __launch_bounds__(256)
l2_normalize_kernel(const float* __restrict__ x, int rows, int D, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const float* xr = x + static_cast<long long>(row) * D;
  float s = 0.f;
  for (int i = lane; i < D; i += 32) s += xr[i] * xr[i];
  s = warp_sum(s);
  const float inv = 1.f / fmaxf(sqrtf(s), 1e-12f);
  for (int i = lane; i < D; i += 32) out[static_cast<long long>(row) * D + i] = xr[i] * inv;
}
cudaError_t launch_l2_normalize(const float* x, int rows, int D, float* out, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  l2_normalize_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, rows, D, out);
  return cudaGetLastError();
}

// =================================================================================================
// similarity tail (clip.rs:102-121, 144-163, 174-185): warp-per-row dot products, fused multiply-add with
// logit scale/bias, then sigmoid per logit or a max-subtracted softmax over all N logits.
// =================================================================================================
__global__ void __launch_bounds__(256)
similarity_logits_kernel(const float* __restrict__ A, const float* __restrict__ b, int N, int D, float scale, float bias,
                         int activation, float* __restrict__ probs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= N) return;
  const float* a = A + static_cast<long long>(row) * D;
  float s = 0.f;
  for (int i = lane; i < D; i += 32) s = fmaf(a[i], __ldg(b + i), s);
  s = warp_sum(s);
  if (lane == 0) {
    const float logit = fmaf(s, scale, bias);  // f32::mul_add
    probs[row] = activation == 1 ? 1.0f / (1.0f + expf(-logit)) : logit;
  }
}
__global__ void __launch_bounds__(1024) softmax_inplace_kernel(float* __restrict__ p, int N) {
  __shared__ float red[32];
  __shared__ float bc;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float m = -INFINITY;
  for (int i = tid; i < N; i += 1024) m = fmaxf(m, p[i]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (warp == 0) {
    float v = red[lane];
    v = warp_max(v);
    if (lane == 0) bc = v;
  }
  __syncthreads();
  const float gmax = bc;
  float s = 0.f;
  for (int i = tid; i < N; i += 1024) {
    const float e = expf(p[i] - gmax);
    p[i] = e;
    s += e;
  }
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (warp == 0) {
    float v = red[lane];
    v = warp_sum(v);
    if (lane == 0) bc = v;
  }
  __syncthreads();
  const float total = bc;
  for (int i = tid; i < N; i += 1024) p[i] = p[i] / total;
}
cudaError_t launch_similarity(const float* A, const float* b, int N, int D, float scale, float bias, int activation,
                              float* probs, float* /*scratch*/, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  similarity_logits_kernel<<<(N + 7) / 8, 256, 0, st>>>(A, b, N, D, scale, bias, activation, probs);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (activation == 0) {  // 1 = sigmoid (done above), 2 = raw logits (Clip::compare, clip.rs:81-90)
    softmax_inplace_kernel<<<1, 1024, 0, st>>>(probs, N);
    e = cudaGetLastError();
  }
  return e;
}

// =================================================================================================
// weight conversion (load time)
// =================================================================================================
__global__ void convert_f32_bf16_kernel(const float* __restrict__ src, int rows, int cols, int ld_out, int transpose,
                                        __nv_bfloat16* __restrict__ dst) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(rows) * ld_out;
  if (idx >= total) return;
  const int r = static_cast<int>(idx / ld_out), c = static_cast<int>(idx - static_cast<long long>(r) * ld_out);
  float v = 0.f;
  if (c < cols) v = transpose ? src[static_cast<long long>(c) * rows + r] : src[static_cast<long long>(r) * cols + c];
  dst[idx] = __float2bfloat16(v);
}
cudaError_t launch_convert_f32_bf16(const float* src, int rows, int cols, int ld_out, bool transpose,
                                    __nv_bfloat16* dst, cudaStream_t st) {
  const long long total = static_cast<long long>(rows) * ld_out;
  if (total <= 0) return cudaSuccess;
  convert_f32_bf16_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(src, rows, cols, ld_out,
                                                                                      transpose ? 1 : 0, dst);
  return cudaGetLastError();
}

}  // namespace clipb200
