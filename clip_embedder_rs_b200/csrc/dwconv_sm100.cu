// Stride-1 depthwise KxK convolution (K = 3 | 7) for the FastViT / MobileCLIP2 trunk, NHWC fp32 activations:
// the RepMixer token mixer, the 7x7 depthwise conv in front of every ConvMlp and the RepCPE positional encoding
// (83 of the 91 depthwise launches of one MobileCLIP2-S2 forward; SURVEY.md Appendix A "C2").  They replace the
// grouped `Conv` nodes onnxruntime executes for the re-parameterised graph (reference pull_onnx.py:110-116).
//
// B200 design: the halo tile is fetched by ONE TMA instruction.  A 4-D tensor map over [n][H][W][C] with box
// {32 channels, 16+K-1, 16+K-1, 1} lands a dense [y][x][c] tile in shared memory; coordinates that fall outside the
// image (the conv's zero padding) or beyond C (channel tails such as 80 = 32+32+16) are zero-filled by the TMA unit, so
// the kernel has no address arithmetic, no bounds checks and no load instructions for the input at all.  While the copy
// is in flight every thread pulls its K*K taps into registers.  Lanes are channels (conflict-free shared-memory reads,
// 128-byte coalesced stores); each warp owns two output rows of 16 pixels and keeps 2x16 accumulators in registers, so
// one staged input row is read once and feeds both output rows (5.5 LDS per 49 FFMA for K = 7).
#include <cuda.h>

#include "conv_kernels.cuh"
#include "tensormap.h"  // get_encode_tiled
#include "ptx_sm100.cuh"

namespace clipb200 {

namespace {

constexpr int DW_TW = 16, DW_TH = 16, DW_CI = 32, DW_THREADS = 256;

__device__ __forceinline__ void tma_load_4d(const void* tmap, uint64_t* bar, void* smem_dst, int32_t c0, int32_t c1,
                                            int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
      : "memory");
}

// (A packed-f32x2 variant of the 7x7 path — pairs of neighbouring output pixels per FFMA2 — was measured neutral in round 2:
// depthwise class 8.62 ms scalar vs 8.76 ms packed, profiles/r02a_bench.json vs r02g_bench.json; an FFMA2 occupies the FMA
// pipe for two cycles, so only issue slots are saved.  It lives in the git history)
__device__ __forceinline__ void store_out(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_out(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

// CI = channels per CTA block: 32 (one lane per channel, one tile per iteration) or 16 — for widths like 80 = 5 x 16,
// where blocks of 32 leave the lanes of the last block half idle (a sixth of the first S2 stage's work): the two
// half-warps then work on TWO consecutive tiles of the same 16 channels (lane = 16 * tile-in-pair + channel), each staged
// by its own TMA box (the two half-warps then share banks: two-way conflicts on the 22 shared-memory reads per staged
// row, which the FFMA-bound loop hides).
template <int K, typename Tout, int CI>
__global__ void __launch_bounds__(DW_THREADS, 2)
dwconv_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const float* __restrict__ w /*[K*K][C]*/,
                  const float* __restrict__ bias, Tout* __restrict__ out, int H, int W, int C, int tiles_x,
                  int tiles_per_img, int total_tiles) {
  constexpr int IW = DW_TW + K - 1, IH = DW_TH + K - 1;
  constexpr int NT = 32 / CI;                        // tiles per iteration
  constexpr int SUB = IH * IW * CI;                  // floats between the staged tiles of a pair (a multiple of 128 B)
  extern __shared__ __align__(128) float tile[];  // [NT][IH][IW][CI]; declared aligned so the reads stay LDS (no generic LD)
  __shared__ __align__(8) uint64_t bar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.y * CI;
  const int ch = lane % CI, sub = lane / CI;
  const int iters = (total_tiles + NT - 1) / NT;
  // Persistent over the (image, tile) pairs of one channel block: the K*K taps are fetched ONCE per CTA instead of
  // once per tile (the per-tile prologue — 49 predicated LDGs with 64-bit address arithmetic, barrier set-up, CTA launch —
  // was 12 % of the instruction stream of a kernel that is bound by instruction issue), and the next tile's TMA load is
  // requested as soon as the whole CTA has finished reading the current one, i.e. under the output stores; the second
  // CTA resident on the SM computes meanwhile.
  auto issue_load = [&](int it) {
    ptx::mbar_arrive_expect_tx(&bar, NT * IH * IW * CI * 4);
#pragma unroll
    for (int s = 0; s < NT; ++s) {
      const int t = it * NT + s;   // a tile index beyond the last one lands on an image beyond n: zeros, never stored
      const int b = t / tiles_per_img, r = t - b * tiles_per_img;
      tma_load_4d(&tm_in, &bar, tile + s * SUB, c0, (r % tiles_x) * DW_TW - K / 2, (r / tiles_x) * DW_TH - K / 2, b);
    }
  };
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
    if (static_cast<int>(blockIdx.x) < iters) issue_load(blockIdx.x);
  }
  const int c = c0 + ch;
  const bool c_ok = c < C;
  float wk[K * K];
  {
    const float* wp = w + c;
#pragma unroll
    for (int t = 0; t < K * K; ++t, wp += C) wk[t] = c_ok ? __ldg(wp) : 0.f;
  }
  const float bv = c_ok ? __ldg(bias + c) : 0.f;
  __syncthreads();  // the barrier init is visible to everyone
  const int r0 = warp * 2;  // this warp's two output rows inside the tile
  uint32_t phase = 0;
  for (int it = blockIdx.x; it < iters; it += gridDim.x, phase ^= 1) {
    const int t = it * NT + sub;
    const int b = t / tiles_per_img, r = t - b * tiles_per_img;
    const int ty0 = (r / tiles_x) * DW_TH, tx0 = (r % tiles_x) * DW_TW;
    ptx::mbar_wait(&bar, phase);
    float acc0[DW_TW], acc1[DW_TW];
#pragma unroll
    for (int x = 0; x < DW_TW; ++x) acc0[x] = acc1[x] = bv;
#pragma unroll
    for (int iy = 0; iy < K + 1; ++iy) {  // staged rows r0 .. r0+K feed output rows r0 (taps ky = iy) and r0+1 (ky = iy-1)
      float rv[IW];
      const float* src = tile + sub * SUB + ((r0 + iy) * IW) * CI + ch;
#pragma unroll
      for (int i = 0; i < IW; ++i) rv[i] = src[i * CI];
      if (iy < K) {
#pragma unroll
        for (int x = 0; x < DW_TW; ++x)
#pragma unroll
          for (int kx = 0; kx < K; ++kx) acc0[x] = fmaf(rv[x + kx], wk[iy * K + kx], acc0[x]);
      }
      if (iy > 0) {
#pragma unroll
        for (int x = 0; x < DW_TW; ++x)
#pragma unroll
          for (int kx = 0; kx < K; ++kx) acc1[x] = fmaf(rv[x + kx], wk[(iy - 1) * K + kx], acc1[x]);
      }
    }
    __syncthreads();  // every warp has read its rows: the tile may be overwritten
    if (threadIdx.x == 0 && it + static_cast<int>(gridDim.x) < iters) issue_load(it + gridDim.x);
    if (c_ok && t < total_tiles) {
      const int oy = ty0 + r0;
      Tout* o0 = out + ((static_cast<long long>(b) * H + oy) * W + tx0) * C + c;
      Tout* o1 = o0 + static_cast<long long>(W) * C;
      if (ty0 + DW_TH <= H && tx0 + DW_TW <= W) {   // interior tile: no per-pixel bounds checks
#pragma unroll
        for (int x = 0; x < DW_TW; ++x, o0 += C, o1 += C) {
          store_out(o0, acc0[x]);
          store_out(o1, acc1[x]);
        }
      } else {
#pragma unroll
        for (int x = 0; x < DW_TW; ++x, o0 += C, o1 += C) {
          if (tx0 + x < W) {
            if (oy < H) store_out(o0, acc0[x]);
            if (oy + 1 < H) store_out(o1, acc1[x]);
          }
        }
      }
    }
  }
}

// Small feature maps (H, W <= TE, TE = 8 | 4: the last FastViT stages at 8x8 and 4x4): with 16x16 tiles three quarters
// (fifteen sixteenths) of every tile's FMAs multiplied padding — the 8x8 stage of MobileCLIP2-S2 took twice as long per
// launch as the 16x16 stage with twice the elements.  Here one CTA iteration takes 16 / TE whole images: the TMA box
// spans {32 channels, TE+K-1, TE+K-1, 16/TE images}, every warp still owns two output rows (of TE pixels) and the
// persistent / taps-once structure is the one above.
template <int K, typename Tout, int TE>
__global__ void __launch_bounds__(DW_THREADS, 2)
dwconv_small_kernel(const __grid_constant__ CUtensorMap tm_in, const float* __restrict__ w /*[K*K][C]*/,
                    const float* __restrict__ bias, Tout* __restrict__ out, int H, int W, int C, int n, int groups) {
  constexpr int NI = 16 / TE, IE = TE + K - 1;   // images per iteration, staged tile edge
  extern __shared__ __align__(128) float tile[];  // [NI][IE][IE][32]
  __shared__ __align__(8) uint64_t bar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.y * DW_CI;
  auto issue_load = [&](int g) {
    ptx::mbar_arrive_expect_tx(&bar, NI * IE * IE * DW_CI * 4);
    tma_load_4d(&tm_in, &bar, tile, c0, -(K / 2), -(K / 2), g * NI);   // images beyond n arrive as zeros
  };
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
    if (static_cast<int>(blockIdx.x) < groups) issue_load(blockIdx.x);
  }
  const int c = c0 + lane;
  const bool c_ok = c < C;
  float wk[K * K];
  {
    const float* wp = w + c;
#pragma unroll
    for (int t = 0; t < K * K; ++t, wp += C) wk[t] = c_ok ? __ldg(wp) : 0.f;
  }
  const float bv = c_ok ? __ldg(bias + c) : 0.f;
  __syncthreads();
  const int slot = (2 * warp) / TE, r0 = (2 * warp) % TE;   // which image of the group, first of this warp's two rows
  uint32_t phase = 0;
  for (int g = blockIdx.x; g < groups; g += gridDim.x, phase ^= 1) {
    ptx::mbar_wait(&bar, phase);
    float acc0[TE], acc1[TE];
#pragma unroll
    for (int x = 0; x < TE; ++x) acc0[x] = acc1[x] = bv;
#pragma unroll
    for (int iy = 0; iy < K + 1; ++iy) {
      float rv[IE];
      const float* src = tile + ((slot * IE + r0 + iy) * IE) * DW_CI + lane;
#pragma unroll
      for (int i = 0; i < IE; ++i) rv[i] = src[i * DW_CI];
      if (iy < K) {
#pragma unroll
        for (int x = 0; x < TE; ++x)
#pragma unroll
          for (int kx = 0; kx < K; ++kx) acc0[x] = fmaf(rv[x + kx], wk[iy * K + kx], acc0[x]);
      }
      if (iy > 0) {
#pragma unroll
        for (int x = 0; x < TE; ++x)
#pragma unroll
          for (int kx = 0; kx < K; ++kx) acc1[x] = fmaf(rv[x + kx], wk[(iy - 1) * K + kx], acc1[x]);
      }
    }
    __syncthreads();  // every warp has read its rows: the tile may be overwritten
    if (threadIdx.x == 0 && g + static_cast<int>(gridDim.x) < groups) issue_load(g + gridDim.x);
    const int b = g * NI + slot;
    if (c_ok && b < n) {
      Tout* o0 = out + ((static_cast<long long>(b) * H + r0) * W) * C + c;
      Tout* o1 = o0 + static_cast<long long>(W) * C;
#pragma unroll
      for (int x = 0; x < TE; ++x, o0 += C, o1 += C) {
        if (x < W) {
          if (r0 < H) store_out(o0, acc0[x]);
          if (r0 + 1 < H) store_out(o1, acc1[x]);
        }
      }
    }
  }
}

bool make_tmap_nhwc_f32(CUtensorMap* tm, const float* base, int n, int H, int W, int C, int box_w, int box_h, int box_c = DW_CI) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return false;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(n)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 4, static_cast<cuuint64_t>(W) * C * 4,
                           static_cast<cuuint64_t>(H) * W * C * 4};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_tmap_nhwc_f32_small(CUtensorMap* tm, const float* base, int n, int H, int W, int C, int box_e, int box_n) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return false;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(n)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 4, static_cast<cuuint64_t>(W) * C * 4,
                           static_cast<cuuint64_t>(H) * W * C * 4};
  cuuint32_t box[4] = {DW_CI, static_cast<cuuint32_t>(box_e), static_cast<cuuint32_t>(box_e), static_cast<cuuint32_t>(box_n)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline int resident_ctas() {
  static const int resident = [] {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return 2 * sms;
  }();
  return resident;
}

template <int K, typename Tout, int TE>
cudaError_t launch_small_t(const float* in, int n, int H, int W, int C, const float* w, const float* bias, Tout* out,
                           cudaStream_t st) {
  constexpr int NI = 16 / TE, IE = TE + K - 1;
  constexpr int smem = NI * IE * IE * DW_CI * 4;
  CUtensorMap tm;
  if (!make_tmap_nhwc_f32_small(&tm, in, n, H, W, C, IE, NI)) return cudaErrorInvalidValue;
  const int groups = (n + NI - 1) / NI, cblocks = (C + DW_CI - 1) / DW_CI;
  int per_block = resident_ctas() / cblocks;
  if (per_block < 1) per_block = 1;
  if (per_block > groups) per_block = groups;
  dwconv_small_kernel<K, Tout, TE><<<dim3(per_block, cblocks, 1), DW_THREADS, smem, st>>>(tm, w, bias, out, H, W, C, n, groups);
  return cudaGetLastError();
}

template <int K, typename Tout>
cudaError_t launch_t(const float* in, int n, int H, int W, int C, const float* w, const float* bias, Tout* out,
                     cudaStream_t st) {
  static const bool small_on = [] { const char* v = getenv("CLIPB200_DWCONV_SMALL"); return v == nullptr || atoi(v) != 0; }();
  if (small_on && H <= 4 && W <= 4) return launch_small_t<K, Tout, 4>(in, n, H, W, C, w, bias, out, st);
  if (small_on && H <= 8 && W <= 8) return launch_small_t<K, Tout, 8>(in, n, H, W, C, w, bias, out, st);
  constexpr int IW = DW_TW + K - 1, IH = DW_TH + K - 1;
  constexpr int smem = IH * IW * DW_CI * 4;
  static const bool half_on = [] { const char* v = getenv("CLIPB200_DWCONV_HALF_BLOCKS"); return v == nullptr || atoi(v) != 0; }();
  const bool half = half_on && C % 32 == 16;   // e.g. 80 channels: five blocks of 16 (two tiles per iteration) instead of 32 + 32 + 16
  const int ci = half ? 16 : DW_CI, nt = 32 / ci;
  CUtensorMap tm;
  if (!make_tmap_nhwc_f32(&tm, in, n, H, W, C, IW, IH, ci)) return cudaErrorInvalidValue;
  const int tiles_x = (W + DW_TW - 1) / DW_TW, tiles_y = (H + DW_TH - 1) / DW_TH;
  const int tiles_per_img = tiles_x * tiles_y, total = tiles_per_img * n, cblocks = (C + ci - 1) / ci;
  const int iters = (total + nt - 1) / nt;
  // persistent CTAs: two per SM over all channel blocks together
  int per_block = resident_ctas() / cblocks;
  if (per_block < 1) per_block = 1;
  if (per_block > iters) per_block = iters;
  dim3 grid(per_block, cblocks, 1);
  if (half) dwconv_tma_kernel<K, Tout, 16><<<grid, DW_THREADS, smem, st>>>(tm, w, bias, out, H, W, C, tiles_x, tiles_per_img, total);
  else dwconv_tma_kernel<K, Tout, 32><<<grid, DW_THREADS, smem, st>>>(tm, w, bias, out, H, W, C, tiles_x, tiles_per_img, total);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Strided and / or channel-multiplying variants (stem depthwise 3x3 s2 on bf16, the 7x7 s2 x2 downsampling convs, the
// final 3x3 x2 expansion): same TMA-staged halo tile, one output row of 16 pixels per warp.  Lanes are OUTPUT channels;
// with a channel multiplier of 2 two neighbouring lanes read the same input channel (a shared-memory broadcast), so the
// staged tile only holds 32 / MULT input channels.
template <int K, int STRIDE, int MULT, typename Tin, typename Tout, bool GELU>
struct DwGen {
  static constexpr int TW = 16, TH = 8, CI = 32 / MULT;
  static constexpr int IW = (TW - 1) * STRIDE + K, IH = (TH - 1) * STRIDE + K;
  static constexpr int SMEM = IH * IW * CI * static_cast<int>(sizeof(Tin));
};

__device__ __forceinline__ float load_in(const float* p) { return *p; }
__device__ __forceinline__ float load_in(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float gelu_erf_dw(float x) {  // Abramowitz-Stegun form, same as gemm_sm100.cuh::gelu_erf_fast
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

template <int K, int STRIDE, int MULT, typename Tin, typename Tout, bool GELU>
__global__ void __launch_bounds__(DW_THREADS, 2)
dwconv_tma_gen_kernel(const __grid_constant__ CUtensorMap tm_in, const float* __restrict__ w /*[K*K][Cout]*/,
                      const float* __restrict__ bias, Tout* __restrict__ out, int Ho, int Wo, int Cout, int tiles_x) {
  using G = DwGen<K, STRIDE, MULT, Tin, Tout, GELU>;
  extern __shared__ __align__(128) uint8_t gen_tile_raw[];
  const Tin* tile = reinterpret_cast<const Tin*>(gen_tile_raw);  // [IH][IW][CI]
  __shared__ __align__(8) uint64_t bar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.y * 32, b = blockIdx.z;  // first OUTPUT channel of this block
  const int ty0 = (blockIdx.x / tiles_x) * G::TH, tx0 = (blockIdx.x % tiles_x) * G::TW;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
    ptx::mbar_arrive_expect_tx(&bar, G::SMEM);
    tma_load_4d(&tm_in, &bar, gen_tile_raw, c0 / MULT, tx0 * STRIDE - K / 2, ty0 * STRIDE - K / 2, b);
  }
  const int oc = c0 + lane;
  const bool c_ok = oc < Cout;
  float wk[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) wk[t] = c_ok ? __ldg(w + t * Cout + oc) : 0.f;
  const float bv = c_ok ? __ldg(bias + oc) : 0.f;
  __syncthreads();
  ptx::mbar_wait(&bar, 0);
  const int cl = lane / MULT;  // input channel inside the staged tile
  float acc[G::TW];
#pragma unroll
  for (int x = 0; x < G::TW; ++x) acc[x] = bv;
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
    float rv[G::IW];
    const Tin* src = tile + ((warp * STRIDE + ky) * G::IW) * G::CI + cl;
#pragma unroll
    for (int i = 0; i < G::IW; ++i) rv[i] = load_in(src + i * G::CI);
#pragma unroll
    for (int x = 0; x < G::TW; ++x)
#pragma unroll
      for (int kx = 0; kx < K; ++kx) acc[x] = fmaf(rv[x * STRIDE + kx], wk[ky * K + kx], acc[x]);
  }
  const int oy = ty0 + warp;
  if (!c_ok || oy >= Ho) return;
  Tout* o0 = out + ((static_cast<long long>(b) * Ho + oy) * Wo + tx0) * Cout + oc;
#pragma unroll
  for (int x = 0; x < G::TW; ++x)
    if (tx0 + x < Wo) store_out(o0 + static_cast<long long>(x) * Cout, GELU ? gelu_erf_dw(acc[x]) : acc[x]);
}

template <typename Tin>
bool make_tmap_nhwc(CUtensorMap* tm, const Tin* base, int n, int H, int W, int C, int box_c, int box_w, int box_h) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return false;
  const cuuint64_t es = sizeof(Tin);
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(n)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * es, static_cast<cuuint64_t>(W) * C * es,
                           static_cast<cuuint64_t>(H) * W * C * es};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(tm, sizeof(Tin) == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
             const_cast<Tin*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int K, int STRIDE, int MULT, typename Tin, typename Tout, bool GELU>
cudaError_t launch_gen(const Tin* in, int n, int H, int W, int C, const float* w, const float* bias, Tout* out,
                       cudaStream_t st) {
  using G = DwGen<K, STRIDE, MULT, Tin, Tout, GELU>;
  const int Ho = (H + 2 * (K / 2) - K) / STRIDE + 1, Wo = (W + 2 * (K / 2) - K) / STRIDE + 1, Cout = C * MULT;
  CUtensorMap tm;
  if (!make_tmap_nhwc<Tin>(&tm, in, n, H, W, C, G::CI, G::IW, G::IH)) return cudaErrorInvalidValue;
  const int tiles_x = (Wo + G::TW - 1) / G::TW, tiles_y = (Ho + G::TH - 1) / G::TH;
  dim3 grid(tiles_x * tiles_y, (Cout + 31) / 32, n);
  dwconv_tma_gen_kernel<K, STRIDE, MULT, Tin, Tout, GELU><<<grid, DW_THREADS, G::SMEM, st>>>(tm, w, bias, out, Ho, Wo, Cout, tiles_x);
  return cudaGetLastError();
}

template <int K, int STRIDE, int MULT, typename Tin, typename Tout, bool GELU>
cudaError_t configure_gen() {
  return cudaFuncSetAttribute(dwconv_tma_gen_kernel<K, STRIDE, MULT, Tin, Tout, GELU>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, DwGen<K, STRIDE, MULT, Tin, Tout, GELU>::SMEM);
}

template <int K, typename Tout>
cudaError_t configure_t() {
  constexpr int smem = (DW_TH + K - 1) * (DW_TW + K - 1) * DW_CI * 4;
  cudaError_t e = cudaFuncSetAttribute(dwconv_tma_kernel<K, Tout, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(dwconv_tma_kernel<K, Tout, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(dwconv_small_kernel<K, Tout, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             2 * (8 + K - 1) * (8 + K - 1) * DW_CI * 4);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(dwconv_small_kernel<K, Tout, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             4 * (4 + K - 1) * (4 + K - 1) * DW_CI * 4);
  return e;
}

}  // namespace

// opt in to > 48 KB of dynamic shared memory once per device (called from Engine::Init, never inside a graph capture)
cudaError_t dwconv_tma_configure_device() {
  cudaError_t e = configure_t<7, __nv_bfloat16>();
  if (e == cudaSuccess) e = configure_t<7, float>();
  if (e == cudaSuccess) e = configure_t<3, __nv_bfloat16>();
  if (e == cudaSuccess) e = configure_t<3, float>();
  if (e == cudaSuccess) e = configure_gen<3, 2, 1, __nv_bfloat16, __nv_bfloat16, true>();
  if (e == cudaSuccess) e = configure_gen<7, 2, 2, float, __nv_bfloat16, true>();
  if (e == cudaSuccess) e = configure_gen<7, 2, 2, float, float, false>();
  if (e == cudaSuccess) e = configure_gen<3, 1, 2, float, float, false>();
  return e;
}

// The strided / multiplier combinations of the FastViT trunk; false = not one of them (caller falls back).
bool launch_dwconv_tma_gen(const void* in, bool in_bf16, int n, int H, int W, int C, int K, int stride, int mult,
                           const float* w, const float* bias, bool gelu, void* out, bool out_bf16, cudaStream_t st,
                           cudaError_t* err) {
  const size_t es = in_bf16 ? 2 : 4;
  if ((C * es) % 16 != 0 || (H % stride) != 0 || (W % stride) != 0) return false;  // TMA stride rule; even sizes
  if (K == 3 && stride == 2 && mult == 1 && in_bf16 && out_bf16 && gelu) {
    *err = launch_gen<3, 2, 1, __nv_bfloat16, __nv_bfloat16, true>(static_cast<const __nv_bfloat16*>(in), n, H, W, C, w, bias,
                                                                    static_cast<__nv_bfloat16*>(out), st);
    return true;
  }
  if (K == 7 && stride == 2 && mult == 2 && !in_bf16 && out_bf16 && gelu) {
    *err = launch_gen<7, 2, 2, float, __nv_bfloat16, true>(static_cast<const float*>(in), n, H, W, C, w, bias,
                                                           static_cast<__nv_bfloat16*>(out), st);
    return true;
  }
  if (K == 7 && stride == 2 && mult == 2 && !in_bf16 && !out_bf16 && !gelu) {
    *err = launch_gen<7, 2, 2, float, float, false>(static_cast<const float*>(in), n, H, W, C, w, bias, static_cast<float*>(out), st);
    return true;
  }
  if (K == 3 && stride == 1 && mult == 2 && !in_bf16 && !out_bf16 && !gelu) {
    *err = launch_gen<3, 1, 2, float, float, false>(static_cast<const float*>(in), n, H, W, C, w, bias, static_cast<float*>(out), st);
    return true;
  }
  return false;
}

bool dwconv_tma_supported(bool in_bf16, int C, int K, int stride, int mult, bool gelu) {
  // TMA needs 16-byte global strides: C * 4 bytes per pixel -> C % 4 == 0
  return !in_bf16 && stride == 1 && mult == 1 && !gelu && (K == 3 || K == 7) && C % 4 == 0;
}

cudaError_t launch_dwconv_tma(const float* in, int n, int H, int W, int C, int K, const float* w, const float* bias,
                              void* out, bool out_bf16, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (K == 7 && out_bf16) return launch_t<7, __nv_bfloat16>(in, n, H, W, C, w, bias, static_cast<__nv_bfloat16*>(out), st);
  if (K == 7) return launch_t<7, float>(in, n, H, W, C, w, bias, static_cast<float*>(out), st);
  if (K == 3 && out_bf16) return launch_t<3, __nv_bfloat16>(in, n, H, W, C, w, bias, static_cast<__nv_bfloat16*>(out), st);
  if (K == 3) return launch_t<3, float>(in, n, H, W, C, w, bias, static_cast<float*>(out), st);
  return cudaErrorInvalidValue;
}

}  // namespace clipb200
