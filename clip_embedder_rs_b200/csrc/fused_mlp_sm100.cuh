// Fused ConvMlp tail of the FastViT / MobileCLIP2 blocks for sm_100a:
//
//     x[r, :] += gamma ⊙ ( W2 · gelu( W1 · a[r, :] + b1 ) + b2 )          a = dw7x7(x) (bf16, from dwconv_sm100.cu)
//
// i.e. the two 1x1 convolutions of timm's ConvMlp with the GELU between them, the layer scale and the residual add
// (reference: the Conv -> Erf-GELU -> Conv -> Mul -> Add nodes onnxruntime executes for every re-parameterised RepMixer /
// attention block, pull_onnx.py:110-116).  As two GEMM launches the 3C-wide hidden activation crosses HBM twice (6C
// bytes written by fc1, 6C read by fc2 per pixel, C = 80 ... 320), which bounds fc1 at 60 ... 240 FLOP/B — 390 ... 1560
// TFLOP/s at the copy peak — and is why the 1x1-conv GEMM class sat at 404 TFLOP/s.  Here the hidden
// activation never leaves the SM:
//
//   persistent CTAs, one tile of 128 pixels at a time.  A [128 x C] is loaded once (TMA, 128B swizzle, K-major).  The hidden dimension is
//   walked in chunks of 64 columns:
//       MMA warp   S_b[128 x 64]  = A · W1[chunk]^T          tcgen05.mma, accumulator in TMEM buffer b (2 buffers)
//       4 warps    h = gelu(S_b + b1) -> bf16 -> shared memory tile H_b [128 x 64] written in the 128B-swizzled K-major
//                  layout the tensor core reads (2 buffers)
//       MMA warp   O[128 x C]    += H_b · W2[:, chunk]^T      accumulator in TMEM (C columns)
//   W1 / W2 chunks stream through TMA rings; chunk i's GELU overlaps chunk i+1's first GEMM and chunk i-1's second.
//   Epilogue: O -> (+ b2) * gamma -> staging -> TMA reduce-add into the fp32 residual stream (as the GEMM's EPI_RESID).
//
// HBM traffic per pixel: 2C (a) + 8C (x read-modify-write) instead of 2C + 6C + 6C + 8C.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_sm100.cuh"
#include "ptx_sm100.cuh"

namespace clipb200 {
namespace fmlp {

// Build-time A/B switches; the defaults are what shipped, the alternatives were measured on one box and lost
// (profiles/r02p_fused_mlp_ab.md):
//   CLIPB200_FMLP_GROUPS      1: all 16 epilogue warps work on every hidden chunk;  2: two groups of 8 take alternate
//                                chunks (one S / H buffer each) so that one group's GELU overlaps the other's TMEM load,
//                                shared-memory store, proxy fence and barrier round trip (C = 160: 162 -> 179 us)
//   CLIPB200_FMLP_DIRECT_RED  0: staging tile + TMA reduce-add into x;  1: red.global.add.v4.f32 straight from the
//                                registers, every thread owning 64 contiguous bytes of its pixel row (C = 80: 271 -> 302 us)
//   CLIPB200_FMLP_POLL        0: the MMA warp issues S(i + 1) then O += H(i) in a fixed interleaving;  1: it polls both
//                                streams and issues whichever is ready, S up to two chunks ahead (C = 320: 118 -> 136 us)
#ifndef CLIPB200_FMLP_GROUPS
#define CLIPB200_FMLP_GROUPS 1
#endif
#ifndef CLIPB200_FMLP_DIRECT_RED
#define CLIPB200_FMLP_DIRECT_RED 0
#endif
#ifndef CLIPB200_FMLP_POLL
#define CLIPB200_FMLP_POLL 0
#endif
//   CLIPB200_FMLP_CLUSTER     0: independent CTAs;  1: CTAs run as clusters of two on neighbouring tiles and share every W1 /
//                                W2 tile (each CTA issues half of the TMA loads, multicast into both CTAs' rings; both MMA
//                                warps release a slot), so the weights cross the L2 -> SM path once per PAIR of tiles.
//                                Correct on every case, and no faster: C = 320 121.8 vs 119.1 us, C = 160 164.3 vs 164.0,
//                                C = 80 273.1 vs 270.2 (profiles/r02ac_fmlp_cluster.log) — the MMA warp's waits for W1 are
//                                latency and coupling, not L2 bandwidth
#ifndef CLIPB200_FMLP_CLUSTER
#define CLIPB200_FMLP_CLUSTER 0
#endif
//   CLIPB200_FMLP_W_ORDER     0: weight tiles requested as W1(i), W2(i), W1(i + 1), ...;  1: in the MMA warp's consumption
//                                order (W1 one chunk ahead of W2), with CLIPB200_FMLP_NS1_WIDE / _NS2_WIDE = the ring depths of
//                                the C > 256 instances.  Order x {4+4, 6+3, 8+2} stages all land on 117 / 163 / 272 us
//                                (profiles/r02ad_fmlp_worder.log): the MMA warp does wait for W1 a quarter of a C = 320 tile,
//                                but it is not on the critical path — the epilogue warps are
#ifndef CLIPB200_FMLP_W_ORDER
#define CLIPB200_FMLP_W_ORDER 0
#endif
#ifndef CLIPB200_FMLP_NS1_WIDE
#define CLIPB200_FMLP_NS1_WIDE 4
#endif
#ifndef CLIPB200_FMLP_NS2_WIDE
#define CLIPB200_FMLP_NS2_WIDE 4
#endif

constexpr int BM = 128;        // pixels per CTA
constexpr int HC = 64;         // hidden columns per chunk (one 128B-swizzle atom of K for the second GEMM)
// Epilogue warps: SLOTS per TMEM lane quarter, each taking 64 / SLOTS of a chunk's hidden columns.  The GELU between the
// GEMMs (~9 instructions and 2 MUFU ops per hidden element) is what a chunk waits for, so the SM wants 16 of them:
// 8 per CTA where two CTAs share an SM (C <= 96), 16 where one CTA has it alone.  Then the TMA warp, then the MMA warp.

template <int C>
struct Cfg {
  static_assert(C % 16 == 0 && C >= 16 && C <= 320, "channel count");
  static constexpr int KB = (C + 63) / 64;                 // 64-column k-blocks of A / W1
  static constexpr int KSTEPS1 = C / 16;                   // UMMA k-steps of the first GEMM
  static constexpr int A_BYTES = KB * BM * 128;            // [KB][128 rows][128 B]
  // W1 streams in tiles of [64 hidden rows][64 input channels] (one per k-block of a chunk), W2 in tiles of
  // [N2 output channels][64 hidden columns] (one or two per chunk), each through its own ring, in the order the MMA warp
  // consumes them: the rings hold about two chunks' worth, so a chunk's weights arrive while the previous one computes.
  static constexpr int N2 = C > 256 ? C / 2 : C;           // second GEMM: N per instruction (<= 256)
  static constexpr int N2_PARTS = C > 256 ? 2 : 1;
  static constexpr int W1_TILE = HC * 128;                 // 8 KB
  static constexpr int W2_TILE = ((N2 + 7) / 8 * 8 * 128 + 1023) / 1024 * 1024;
  static constexpr int H_BYTES = BM * 128;                 // [128 rows][64 bf16]
  // Two CTAs per SM where the tiles are small enough (C <= 96), so that one CTA's prologue / final epilogue overlaps
  // the other's chunks; otherwise one CTA.
  static constexpr int CTAS_PER_SM = C <= 96 ? 2 : 1;
  static constexpr int EPI_WARPS = CTAS_PER_SM == 2 ? 8 : 16;
  static constexpr int SLOTS = EPI_WARPS / 4;              // warps per lane quarter
  static constexpr int GROUPS = (EPI_WARPS == 16 && CLIPB200_FMLP_GROUPS == 2) ? 2 : 1;   // warp groups on alternate chunks
  static constexpr int CSLOTS = SLOTS / GROUPS;            // warps per lane quarter working on one chunk
  static constexpr int CPW = HC / CSLOTS;                  // hidden columns per warp and chunk (32 | 16)
  static constexpr bool DIRECT = CLIPB200_FMLP_DIRECT_RED != 0;
  static constexpr int OCW = DIRECT ? 16 : HC / SLOTS;     // output columns per final-epilogue chunk (staging: 128 B | 64 B rows)
  static constexpr int THREADS = 32 * (EPI_WARPS + 2);
  static constexpr int WARP_TMA = EPI_WARPS, WARP_MMA = EPI_WARPS + 1;
  // ring depths: what the shared memory allows; at C = 320 the MMA warp waits for W1 a quarter of the time whatever the split
  // (4 + 4 measured best; halving the L2 -> SM weight traffic with cluster multicast changes nothing: CLIPB200_FMLP_CLUSTER)
  static constexpr int NS1 = C <= 80 ? KB + 1 : (C <= 96 ? KB : (C > 256 ? CLIPB200_FMLP_NS1_WIDE : (3 * KB < 8 ? 3 * KB : 8)));   // W1 ring stages
  static constexpr int NS2 = C <= 96 ? 2 : (C > 256 ? CLIPB200_FMLP_NS2_WIDE : (C > 192 ? 2 : 3));                                 // W2 ring stages
  static constexpr int STG_WARP = 32 * OCW * 4;            // per epilogue warp: [32 rows][OCW fp32]
  static constexpr int STG_BYTES = EPI_WARPS * STG_WARP;   // aliases the H buffers
  static constexpr int OFF_A = 0;
  static constexpr int OFF_W1 = OFF_A + A_BYTES;
  static constexpr int OFF_W2 = OFF_W1 + NS1 * W1_TILE;
  static constexpr int OFF_H = OFF_W2 + NS2 * W2_TILE;
  static constexpr int OFF_STG = OFF_H;                    // the final epilogue runs after the last second GEMM retired
  static_assert(STG_BYTES <= 2 * H_BYTES, "staging aliases the two H buffers");
  static constexpr int OFF_BAR = OFF_H + 2 * H_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
  static constexpr int COL_O = 0;                          // O accumulator: C columns
  static constexpr int COL_S = (C + 31) / 32 * 32;         // two hidden buffers of HC columns
  static constexpr int TMEM_NEED = COL_S + 2 * HC;
  static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512);
  static_assert(N2 % 16 == 0, "UMMA N granularity");
  static_assert(NS1 <= 8 && NS2 <= 4, "barrier carve");
  static_assert((SMEM_BYTES + 2048) * CTAS_PER_SM <= 228 * 1024, "shared memory (static + reserved included)");
  static_assert(TMEM_COLS * CTAS_PER_SM <= 512, "tensor memory");
};

struct Params {
  int M, C, Hd;          // rows (pixels), channels, hidden width (fc1.N)
  const float* b1;       // [Hd]
  const float* b2;       // [C]
  const float* gamma;    // [C] or null
  float* x;              // [M, ldx] fp32 residual stream (direct reduce mode)
  long long ldx;
#ifdef CLIPB200_FMLP_TIMING
  unsigned long long* dbg;   // 32 counters written by CTA 0 (tests/native/gemm_test.cu prints them)
#endif
};
#ifdef CLIPB200_FMLP_TIMING
// Where-does-the-time-go build (-DCLIPB200_FMLP_TIMING): CTA 0's first epilogue warp and its MMA warp accumulate clock64
// deltas around every wait and every phase; never compiled into the library.
inline unsigned long long*& timing_buffer() { static unsigned long long* p = nullptr; return p; }
#define FMLP_T(var) const long long var = clock64()
#define FMLP_ACC(i, a, b) tacc[i] += (b) - (a)
#else
#define FMLP_T(var)
#define FMLP_ACC(i, a, b)
#endif

__device__ __forceinline__ float gelu_erf(float x) { return gelu_erf_fast(x); }

// gelu_erf_fast (gemm_sm100.cuh: x * Phi(x), Phi from Abramowitz-Stegun 7.1.26, |error| <= 4.3e-7) for TWO values at
// once in packed f32x2 arithmetic: the polynomial, the scalings and the final x * Phi are one FFMA2 / FMUL2 per pair
// instead of one instruction per element; the two MUFU ops per element (rcp, ex2) remain.  ~8 instead of ~14
// instructions per element — the GELU between the two GEMMs is what the fused kernel's epilogue warps spend their time on.
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_splat(float v) { return f2_pack(v, v); }
// returns bf16x2 {gelu(x0 + b0), gelu(x1 + b1)}
__device__ __forceinline__ uint32_t gelu_erf_pair_bf16(float x0, float x1, float b0, float b1) {
#if defined(CLIPB200_FMLP_EXPERIMENT) && CLIPB200_FMLP_EXPERIMENT == 1   // timing aid: how much of the kernel is the GELU?  (wrong results)
  __nv_bfloat162 r0 = __floats2bfloat162_rn(x0 + b0, x1 + b1);
  return *reinterpret_cast<uint32_t*>(&r0);
#endif
  // gelu(x) = x * Phi(x) = relu(x) - |x| * Phi(-|x|): no sign select, and h = Phi(-|x|) is what the formula yields
  const float xa = x0 + b0, xb = x1 + b1;
  const uint64_t nax = f2_pack(-fabsf(xa), -fabsf(xb));
  float d0, d1, s0, s1;
  f2_unpack(f2_fma(nax, f2_splat(-0.23164189f), f2_splat(1.0f)), d0, d1);        // 1 + p |x| / sqrt 2 folded: p' = 0.3275911 / sqrt 2
  f2_unpack(f2_mul(f2_mul(nax, nax), f2_splat(-0.72134752f)), s0, s1);          // -x^2 / 2 * log2(e)
  float t0, t1, e0, e1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(s0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(s1));
  const uint64_t t = f2_pack(t0, t1), e = f2_pack(e0, e1);
  uint64_t q = f2_fma(t, f2_splat(0.5307027145f), f2_splat(-0.7265760135f));
  q = f2_fma(t, q, f2_splat(0.7107068705f));
  q = f2_fma(t, q, f2_splat(-0.142248368f));
  q = f2_fma(t, q, f2_splat(0.127414796f));
  const uint64_t h = f2_mul(f2_mul(q, t), e);                                    // h = Phi(-|x|)
  float g0, g1;
  f2_unpack(f2_fma(nax, h, f2_pack(fmaxf(xa, 0.f), fmaxf(xb, 0.f))), g0, g1);
  __nv_bfloat162 r = __floats2bfloat162_rn(g0, g1);
  return *reinterpret_cast<uint32_t*>(&r);
}

template <int C>
__global__ void __launch_bounds__(Cfg<C>::THREADS, Cfg<C>::CTAS_PER_SM)
fused_mlp_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w1,
                 const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_x, Params p) {
  using K = Cfg<C>;
  constexpr int EPI_WARPS = K::EPI_WARPS, WARP_TMA = K::WARP_TMA, WARP_MMA = K::WARP_MMA;
  extern __shared__ uint8_t fmlp_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fmlp_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_a = smem + K::OFF_A;
  uint8_t* s_w1 = smem + K::OFF_W1;
  uint8_t* s_w2 = smem + K::OFF_W2;
  uint8_t* s_h = smem + K::OFF_H;
  uint8_t* s_stg = smem + K::OFF_STG;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::OFF_BAR);
  uint64_t* a_full = bars + 0;
  uint64_t* w1_full = bars + 1;    // [NS1 <= 8]
  uint64_t* w1_empty = bars + 9;   // [NS1]
  uint64_t* w2_full = bars + 17;   // [NS2 <= 4]
  uint64_t* w2_empty = bars + 21;  // [NS2]
  uint64_t* s_full = bars + 25;    // [2]  MMA -> epilogue: hidden accumulator b complete
  uint64_t* s_empty = bars + 27;   // [2]  epilogue -> MMA: accumulator b read out
  uint64_t* h_full = bars + 29;    // [2]  epilogue -> MMA: H_b written
  uint64_t* h_empty = bars + 31;   // [2]  MMA -> epilogue: second GEMM of H_b retired
  uint64_t* o_full = bars + 33;    //      MMA -> epilogue: O of this tile complete
  uint64_t* o_empty = bars + 34;   //      epilogue -> MMA: O read out, the next tile may overwrite it
  uint64_t* a_empty = bars + 35;   //      MMA -> producer: every first GEMM of this tile retired, A may be reloaded
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 36);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int n_chunks = (p.Hd + HC - 1) / HC;
  const int n_tiles = (p.M + BM - 1) / BM;
  // Clusters of two: CTA rank r of cluster c walks tiles 2 * (c + k * n_clusters) + r, i.e. blockIdx.x + k * gridDim.x; both
  // CTAs of a pair run the same number of tiles (an odd tile count gives the last pair a phantom tile beyond M: its A rows
  // load as zeros and its output rows are clipped), because they consume the shared weight rings in lockstep.
  constexpr bool CL = CLIPB200_FMLP_CLUSTER != 0;
  const uint32_t rank = CL ? ptx::cluster_ctarank() : 0;
  auto tile_valid = [&](int tile) { return CL ? ((tile & ~1) < n_tiles) : (tile < n_tiles); };

  if (warp == WARP_TMA && lane == 0) {
    ptx::prefetch_tmap(&tm_a);
    ptx::prefetch_tmap(&tm_w1);
    ptx::prefetch_tmap(&tm_w2);
    ptx::prefetch_tmap(&tm_x);
    ptx::mbar_init(a_full, 1);
    ptx::mbar_init(a_empty, 1);
    for (int i = 0; i < K::NS1; ++i) {
      ptx::mbar_init(&w1_full[i], 1);
      ptx::mbar_init(&w1_empty[i], CL ? 2 : 1);   // clusters: the MMA warps of BOTH CTAs release a shared weight slot
    }
    for (int i = 0; i < K::NS2; ++i) {
      ptx::mbar_init(&w2_full[i], 1);
      ptx::mbar_init(&w2_empty[i], CL ? 2 : 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&s_empty[i], 32 * EPI_WARPS / K::GROUPS);
      ptx::mbar_init(&h_full[i], 32 * EPI_WARPS / K::GROUPS);
      ptx::mbar_init(&h_empty[i], 1);
    }
    ptx::mbar_init(o_full, 1);
    ptx::mbar_init(o_empty, 32 * EPI_WARPS);
    ptx::fence_mbar_init();
  }
  if (warp == WARP_MMA) ptx::tmem_alloc<K::TMEM_COLS>(tmem_base_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  if (CL) ptx::cluster_sync();   // the peer's barriers exist before anything is multicast at them
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_ptr, 0);

  // Persistent CTAs: tile = blockIdx.x, += gridDim.x.  All ring / buffer counters run across tiles, so the producer
  // fetches the next tile's A and first weight tiles, and the MMA warp issues the next tile's first GEMMs, while the
  // epilogue warps are still storing the current tile's output.
  if (warp == WARP_TMA) {
    if (lane == 0) {
      uint32_t t1 = 0, t2 = 0, it = 0;
      for (int tile = blockIdx.x; tile_valid(tile); tile += gridDim.x, ++it) {
        ptx::mbar_wait(a_empty, (it & 1) ^ 1);
        // A tile: KB boxes of 64 columns x 128 rows (columns >= C and rows >= M are zero-filled)
        ptx::mbar_arrive_expect_tx(a_full, K::KB * BM * 128);
        for (int kb = 0; kb < K::KB; ++kb) ptx::tma_load_2d(&tm_a, a_full, s_a + kb * BM * 128, kb * 64, tile * BM);
        auto load_w1 = [&](int i) {
          for (int kb = 0; kb < K::KB; ++kb, ++t1) {   // W1 rows [i*64, +64) (hidden units), columns kb*64.. (input channels)
            const int st = t1 % K::NS1;
            ptx::mbar_wait(&w1_empty[st], ((t1 / K::NS1) & 1) ^ 1);
            ptx::mbar_arrive_expect_tx(&w1_full[st], K::W1_TILE);
            if (!CL) ptx::tma_load_2d(&tm_w1, &w1_full[st], s_w1 + st * K::W1_TILE, kb * 64, i * HC);
            else if ((t1 & 1) == rank) ptx::tma_load_2d_multicast(&tm_w1, &w1_full[st], s_w1 + st * K::W1_TILE, kb * 64, i * HC, 3);
          }
        };
        auto load_w2 = [&](int i) {
          for (int part = 0; part < K::N2_PARTS; ++part, ++t2) {   // W2 rows = output channels, columns [i*64, +64) of the hidden dim
            const int st = t2 % K::NS2;
            ptx::mbar_wait(&w2_empty[st], ((t2 / K::NS2) & 1) ^ 1);
            ptx::mbar_arrive_expect_tx(&w2_full[st], K::N2 * 128);
            if (!CL) ptx::tma_load_2d(&tm_w2, &w2_full[st], s_w2 + st * K::W2_TILE, i * HC, part * K::N2);
            else if ((t2 & 1) == rank) ptx::tma_load_2d_multicast(&tm_w2, &w2_full[st], s_w2 + st * K::W2_TILE, i * HC, part * K::N2, 3);
          }
        };
        if (CLIPB200_FMLP_W_ORDER) {
          load_w1(0);
          for (int i = 0; i < n_chunks; ++i) {
            if (i + 1 < n_chunks) load_w1(i + 1);
            load_w2(i);
          }
        } else {
          for (int i = 0; i < n_chunks; ++i) {
            load_w1(i);
            load_w2(i);
          }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    constexpr uint32_t idesc1 = ptx::make_idesc_bf16_f32(BM, HC);
    constexpr uint32_t idesc2 = ptx::make_idesc_bf16_f32(BM, K::N2);
    const uint32_t a_addr = ptx::smem_u32(s_a);
    uint32_t t1 = 0, t2 = 0;      // running W1 / W2 ring tile counters
#ifdef CLIPB200_FMLP_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long m_begin = clock64();
#endif
    // S(g) = A · W1[chunk]^T into TMEM buffer g & 1 (every wait but the weight ring's is done by the caller)
    auto gemm1 = [&](uint32_t g, bool last_of_tile) {
      const int b = g & 1;
      const uint32_t t_s = tmem_base + K::COL_S + static_cast<uint32_t>(b * HC);
#pragma unroll
      for (int kb = 0; kb < K::KB; ++kb, ++t1) {
        const int st = t1 % K::NS1;
        FMLP_T(mw0);
        ptx::mbar_wait(&w1_full[st], (t1 / K::NS1) & 1);
        FMLP_T(mw1);
        FMLP_ACC(2, mw0, mw1);
        ptx::tc_fence_after();
        const uint32_t w_addr = ptx::smem_u32(s_w1 + st * K::W1_TILE);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (kb * 4 + k < K::KSTEPS1) {
            const uint64_t da = ptx::make_kmajor_sw128_desc(a_addr + kb * BM * 128) + static_cast<uint64_t>(2 * k);
            const uint64_t dw = ptx::make_kmajor_sw128_desc(w_addr) + static_cast<uint64_t>(2 * k);
            ptx::umma_bf16_ss_w(t_s, da, dw, idesc1, (kb | k) != 0 ? 1u : 0u);
          }
        }
        if (CL) ptx::umma_commit_multicast_w(&w1_empty[st], 3); else ptx::umma_commit_w(&w1_empty[st]);
      }
      ptx::umma_commit_w(&s_full[b]);
      if (last_of_tile) ptx::umma_commit_w(a_empty);   // every read of this tile's A has been issued
    };
    // O (+)= H(g) · W2[:, chunk]^T; i = chunk index inside the tile (0 starts a new accumulation)
    auto gemm2 = [&](uint32_t g, int i) {
      const int b = g & 1;
      const uint32_t h_addr = ptx::smem_u32(s_h + b * K::H_BYTES);
#pragma unroll
      for (int part = 0; part < K::N2_PARTS; ++part, ++t2) {
        const int st = t2 % K::NS2;
        FMLP_T(mv0);
        ptx::mbar_wait(&w2_full[st], (t2 / K::NS2) & 1);
        FMLP_T(mv1);
        FMLP_ACC(5, mv0, mv1);
        ptx::tc_fence_after();
        const uint32_t w_addr = ptx::smem_u32(s_w2 + st * K::W2_TILE);
#pragma unroll
        for (int k = 0; k < HC / 16; ++k) {
          const uint64_t dh = ptx::make_kmajor_sw128_desc(h_addr) + static_cast<uint64_t>(2 * k);
          const uint64_t dw = ptx::make_kmajor_sw128_desc(w_addr) + static_cast<uint64_t>(2 * k);
          ptx::umma_bf16_ss_w(tmem_base + K::COL_O + static_cast<uint32_t>(part * K::N2), dh, dw, idesc2, (i | k) != 0 ? 1u : 0u);
        }
        if (CL) ptx::umma_commit_multicast_w(&w2_empty[st], 3); else ptx::umma_commit_w(&w2_empty[st]);
      }
      ptx::umma_commit_w(&h_empty[b]);
    };
#if CLIPB200_FMLP_POLL
    // Two instruction streams over this CTA's chunks (running index g across its tiles), issued as their inputs become
    // ready rather than in a fixed interleaving: the first GEMM of chunk g needs S buffer g & 1 read out (chunk g - 2,
    // early in that chunk's GELU) and, at a tile's first chunk, the A tile; the second GEMM of chunk g needs H(g) written
    // and, at a tile's first chunk, the previous tile's O read out.  The older stream goes first when both are ready.
    // S(g + 2) is therefore under way while chunk g's GELU still runs, which is what lets two warp groups on alternate
    // chunks (CLIPB200_FMLP_GROUPS) both find their next S waiting.  The weight rings are consumed in this order too;
    // the producer's order (W1(c), W2(c), W1(c + 1), ...) cannot deadlock it while the W2 ring holds two chunks.
    // g1 - g2 <= 2 keeps that true with two warp groups as well (their S buffers free up early).
    static_assert(K::NS2 >= 2 * K::N2_PARTS, "W2 ring: two chunks' worth");
    static_assert(!CL, "the polling issue order is not combined with clusters");
    const uint32_t my_tiles = (static_cast<uint32_t>(n_tiles) - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const uint32_t total = my_tiles * static_cast<uint32_t>(n_chunks);
    uint32_t g1 = 0, g2 = 0, it1 = 0, it2 = 0;
    int c1 = 0, c2 = 0;           // chunk index inside the tile of each stream
    while (g2 < total) {
      bool did = false;
      if (g2 < g1 && ptx::mbar_test_wait(&h_full[g2 & 1], (g2 >> 1) & 1) &&
          (c2 != 0 || ptx::mbar_test_wait(o_empty, (it2 & 1) ^ 1))) {
        ptx::tc_fence_after();
        gemm2(g2, c2);
        if (++c2 == n_chunks) {
          ptx::umma_commit_w(o_full);
          c2 = 0;
          ++it2;
        }
        ++g2;
        did = true;
      }
      if (g1 < total && g1 - g2 <= 2 && ptx::mbar_test_wait(&s_empty[g1 & 1], ((g1 >> 1) & 1) ^ 1) &&
          (c1 != 0 || ptx::mbar_test_wait(a_full, it1 & 1))) {
        ptx::tc_fence_after();
        gemm1(g1, c1 + 1 == n_chunks);
        if (++c1 == n_chunks) {
          c1 = 0;
          ++it1;
        }
        ++g1;
        did = true;
      }
      if (!did) {   // sleep on the older stream's barrier for a moment instead of burning the epilogue warps' issue slots
        FMLP_T(mi0);
        if (g2 < g1) ptx::mbar_try_wait_hint(&h_full[g2 & 1], (g2 >> 1) & 1, 100u);
        else if (g1 < total) ptx::mbar_try_wait_hint(&s_empty[g1 & 1], ((g1 >> 1) & 1) ^ 1, 100u);
        FMLP_T(mi1);
        FMLP_ACC(3, mi0, mi1);
      }
    }
#else
    uint32_t g1 = 0, g2 = 0, it = 0;
    for (int tile = blockIdx.x; tile_valid(tile); tile += gridDim.x, ++it) {
      ptx::mbar_wait(a_full, it & 1);
      // fixed order: the first GEMM runs one chunk ahead of the second, so chunk i's GELU overlaps S(i + 1)
      auto first = [&](bool last_of_tile) {
        ptx::mbar_wait(&s_empty[g1 & 1], ((g1 >> 1) & 1) ^ 1);
        gemm1(g1, last_of_tile);
        ++g1;
      };
      first(n_chunks == 1);
      for (int i = 0; i < n_chunks; ++i) {
        if (i + 1 < n_chunks) first(i + 2 == n_chunks);
        ptx::mbar_wait(&h_full[g2 & 1], (g2 >> 1) & 1);
        if (i == 0) ptx::mbar_wait(o_empty, (it & 1) ^ 1);   // the previous tile's O has been read out
        gemm2(g2, i);
        ++g2;
      }
      ptx::umma_commit_w(o_full);
    }
#endif
#ifdef CLIPB200_FMLP_TIMING
    if (blockIdx.x == 0 && lane == 0 && p.dbg != nullptr) {
      for (int i = 0; i < 8; ++i) p.dbg[16 + i] = static_cast<unsigned long long>(tacc[i]);
      p.dbg[24] = static_cast<unsigned long long>(clock64() - m_begin);
    }
#endif
  } else {
    // ------------------------------------------------------------ epilogue warps: GELU between the GEMMs, final store
    const int quarter = warp & 3;                              // TMEM lane quarter
    const int slot = warp >> 2;                                // final epilogue: which output column chunks
    const int group = K::GROUPS == 2 ? slot / K::CSLOTS : 0;   // which chunks (running index parity) this warp works on
    const int cslot = slot % K::CSLOTS;                        // which CPW of a chunk's 64 hidden columns
    const int row = quarter * 32 + lane;                       // this thread's pixel row inside the tile
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const int sw = row & 7;                                    // 128B-swizzle phase of this row
    const bool has_b1 = p.b1 != nullptr;
    constexpr int CPW = K::CPW, OCW = K::OCW;
    // Staging tile of this warp, carved out of the H rows of ITS OWN lane quarter (4 KB per quarter in each of the two H
    // buffers = 8 KB for the quarter's SLOTS warps): the only warps that write those H rows are the quarter's own, and
    // they synchronise on the named barrier below before the next tile's first H write.
    uint8_t* stg = s_stg + ((slot * K::STG_WARP) / 4096) * K::H_BYTES + quarter * 4096 + (slot * K::STG_WARP) % 4096;
    static_assert(K::DIRECT || K::SLOTS * K::STG_WARP == 8192, "staging carve");
    (void)stg;
    uint32_t g = 0, it = 0;                                    // running chunk / tile counters
#ifdef CLIPB200_FMLP_TIMING
    long long tacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long e_begin = clock64();
#endif
    for (int tile = blockIdx.x; tile_valid(tile); tile += gridDim.x, ++it) {
      FMLP_T(eb0);
      if (!K::DIRECT && it > 0) {
        // the staging tiles of the previous tile's output alias the H buffers: every TMA read of them must be over, for
        // all warps of this lane quarter (they write interleaved 16-byte chunks of the same H rows)
        if (lane == 0) ptx::tma_store_wait_read<0>();
        asm volatile("bar.sync %0, %1;" ::"r"(quarter + 1), "n"(32 * K::SLOTS) : "memory");
      }
      FMLP_T(eb1);
      FMLP_ACC(7, eb0, eb1);
      for (int i = 0; i < n_chunks; ++i, ++g) {
        const int b = g & 1;
        if (K::GROUPS == 2 && b != group) continue;            // the other group's chunk (it owns S_b and H_b)
        FMLP_T(e0);
        ptx::mbar_wait(&s_full[b], (g >> 1) & 1);
        FMLP_T(e1);
        FMLP_ACC(0, e0, e1);
        ptx::tc_fence_after();
        uint32_t r[CPW];
        if (CPW == 32) ptx::tmem_ld_32x32(tmem_base + lane_base + K::COL_S + static_cast<uint32_t>(b * HC + cslot * CPW), *reinterpret_cast<uint32_t(*)[32]>(r));
        else ptx::tmem_ld_32x32_x16(tmem_base + lane_base + K::COL_S + static_cast<uint32_t>(b * HC + cslot * CPW), *reinterpret_cast<uint32_t(*)[16]>(r));
        const int h0 = i * HC + cslot * CPW;
        float4 bv[CPW / 4];
#pragma unroll
        for (int j = 0; j < CPW / 4; ++j) {
          bv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (has_b1 && h0 + 4 * j < p.Hd) bv[j] = __ldg(reinterpret_cast<const float4*>(p.b1 + h0) + j);   // Hd % 8 == 0
        }
        ptx::tmem_ld_wait();
        FMLP_T(e2);
        FMLP_ACC(1, e1, e2);
        ptx::tc_fence_before();
        ptx::mbar_arrive(&s_empty[b]);
        uint32_t pk[CPW / 2];
#pragma unroll
        for (int j = 0; j < CPW / 4; ++j) {
          pk[2 * j] = gelu_erf_pair_bf16(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), bv[j].x, bv[j].y);
          pk[2 * j + 1] = gelu_erf_pair_bf16(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]), bv[j].z, bv[j].w);
        }
#ifdef CLIPB200_FMLP_TIMING
        asm volatile("" ::"r"(pk[0]), "r"(pk[CPW / 2 - 1]) : "memory");   // the packed values exist before the clock is read
#endif
        FMLP_T(e3);
        FMLP_ACC(2, e2, e3);
        ptx::mbar_wait(&h_empty[b], ((g >> 1) & 1) ^ 1);       // the second GEMM of the chunk two back has finished reading H_b
        FMLP_T(e4);
        FMLP_ACC(3, e3, e4);
        uint8_t* hrow = s_h + b * K::H_BYTES + row * 128;      // 64 bf16 = 8 chunks of 16 B, chunk c stored at (c ^ sw)
#pragma unroll
        for (int c = 0; c < CPW / 8; ++c)
          *reinterpret_cast<uint4*>(hrow + (((cslot * (CPW / 8) + c) ^ sw) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&h_full[b]);
        FMLP_T(e5);
        FMLP_ACC(4, e4, e5);
      }
      FMLP_T(f0);
      // final: O -> (+ b2) * gamma -> staging tile [32 rows][OCW fp32] (swizzled) -> TMA reduce-add into x.  The staging
      // tiles alias the H buffers: o_full completes only after the last second GEMM has finished reading them.
      ptx::mbar_wait(o_full, it & 1);
      FMLP_T(f1);
      FMLP_ACC(5, f0, f1);
      ptx::tc_fence_after();
      constexpr int NCC = (C + OCW - 1) / OCW;                  // OCW-column chunks of O; this warp takes c = slot, slot + SLOTS, ...
      bool released = false;
#pragma unroll 1
      for (int c = slot; c < NCC; c += K::SLOTS) {
        const int n0 = c * OCW;
        uint32_t r[OCW];
        if (OCW == 32) ptx::tmem_ld_32x32(tmem_base + lane_base + K::COL_O + static_cast<uint32_t>(n0), *reinterpret_cast<uint32_t(*)[32]>(r));
        else ptx::tmem_ld_32x32_x16(tmem_base + lane_base + K::COL_O + static_cast<uint32_t>(n0), *reinterpret_cast<uint32_t(*)[16]>(r));
        float4 b4[OCW / 4], g4[OCW / 4];
#pragma unroll
        for (int j = 0; j < OCW / 4; ++j) {
          b4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          g4[j] = make_float4(1.f, 1.f, 1.f, 1.f);
          if (n0 + 4 * j < C) {   // C % 16 == 0
            if (p.b2 != nullptr) b4[j] = __ldg(reinterpret_cast<const float4*>(p.b2 + n0) + j);
            if (p.gamma != nullptr) g4[j] = __ldg(reinterpret_cast<const float4*>(p.gamma + n0) + j);
          }
        }
        ptx::tmem_ld_wait();
        if (c + K::SLOTS >= NCC) {   // this warp's last read of O: the next tile's second GEMM may overwrite it
          ptx::tc_fence_before();
          ptx::mbar_arrive(o_empty);
          released = true;
        }
        if (K::DIRECT) {
          // this thread's 16 consecutive fp32 of its pixel row: four 16-byte reductions, resolved in L2 like the TMA form
          const long long grow = static_cast<long long>(tile) * BM + row;
          if (grow < p.M) {
            float* xr = p.x + grow * p.ldx + n0;
#pragma unroll
            for (int j = 0; j < OCW / 4; ++j)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(xr + 4 * j),
                           "f"((__uint_as_float(r[4 * j]) + b4[j].x) * g4[j].x), "f"((__uint_as_float(r[4 * j + 1]) + b4[j].y) * g4[j].y),
                           "f"((__uint_as_float(r[4 * j + 2]) + b4[j].z) * g4[j].z), "f"((__uint_as_float(r[4 * j + 3]) + b4[j].w) * g4[j].w)
                           : "memory");
          }
        } else {
          if (lane == 0) ptx::tma_store_wait_read<0>();
          __syncwarp();
          // staging rows are 128 B (OCW = 32, 128B swizzle: chunk ^= row & 7) or 64 B (OCW = 16, 64B swizzle: chunk ^= (row >> 1) & 3)
          const int swz = OCW == 32 ? (lane & 7) : ((lane >> 1) & 3);
#pragma unroll
          for (int j = 0; j < OCW / 4; ++j)
            *reinterpret_cast<float4*>(stg + lane * (OCW * 4) + ((j ^ swz) << 4)) =
                make_float4((__uint_as_float(r[4 * j]) + b4[j].x) * g4[j].x, (__uint_as_float(r[4 * j + 1]) + b4[j].y) * g4[j].y,
                            (__uint_as_float(r[4 * j + 2]) + b4[j].z) * g4[j].z, (__uint_as_float(r[4 * j + 3]) + b4[j].w) * g4[j].w);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_reduce_add_2d(&tm_x, stg, n0, tile * BM + quarter * 32);   // columns >= C and rows >= M are clipped
            ptx::tma_store_commit();
          }
        }
      }
      if (!released) {   // a warp without a column chunk of its own (NCC < SLOTS) still has to release O
        ptx::tc_fence_before();
        ptx::mbar_arrive(o_empty);
      }
      FMLP_T(f2);
      FMLP_ACC(6, f1, f2);
    }
#ifdef CLIPB200_FMLP_TIMING
    if (blockIdx.x == 0 && warp == 0 && lane == 0 && p.dbg != nullptr) {
      for (int i = 0; i < 8; ++i) p.dbg[i] = static_cast<unsigned long long>(tacc[i]);
      p.dbg[8] = static_cast<unsigned long long>(clock64() - e_begin);
      p.dbg[9] = it;
      p.dbg[10] = g;
    }
#endif
    if (!K::DIRECT && lane == 0) ptx::tma_store_wait<0>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL) ptx::cluster_sync();   // the peer may still multicast into this CTA's rings / arrive on its barriers until here
  if (warp == WARP_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<K::TMEM_COLS>(tmem_base);
  }
}

template <int C>
inline cudaError_t configure_t() {
  return cudaFuncSetAttribute(fused_mlp_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<C>::SMEM_BYTES);
}

template <int C>
inline cudaError_t launch_t(const __nv_bfloat16* a, long long lda, const __nv_bfloat16* w1, long long ldw1,
                            const __nv_bfloat16* w2, long long ldw2, float* x, long long ldx, const Params& p,
                            int num_sms, cudaStream_t st) {
  using K = Cfg<C>;
  CUtensorMap ta, tw1, tw2, tx;
  if (!make_tmap_2d(&ta, a, p.M, C, lda, BM, 2)) return cudaErrorUnknown;
  if (!make_tmap_2d(&tw1, w1, p.Hd, C, ldw1, HC, 2)) return cudaErrorUnknown;
  if (!make_tmap_2d(&tw2, w2, C, p.Hd, ldw2, K::N2, 2)) return cudaErrorUnknown;   // box = [N2 output channels][64 hidden]
  if (!make_tmap_2d(&tx, x, p.M, C, ldx, 32, 4, K::OCW * 4)) return cudaErrorUnknown;   // box = [32 rows][OCW fp32]
  const int tiles = (p.M + BM - 1) / BM, slots = num_sms * K::CTAS_PER_SM;
  if (CLIPB200_FMLP_CLUSTER) {
    const int pairs = (tiles + 1) / 2, max_clusters = slots / 2;
    const int clusters = pairs < max_clusters ? pairs : max_clusters;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters, 1, 1);
    cfg.blockDim = dim3(K::THREADS, 1, 1);
    cfg.dynamicSmemBytes = K::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, fused_mlp_kernel<C>, ta, tw1, tw2, tx, p);
  }
  const int grid = tiles < slots ? tiles : slots;   // persistent: every CTA walks tiles blockIdx.x, += gridDim.x
  fused_mlp_kernel<C><<<grid, K::THREADS, K::SMEM_BYTES, st>>>(ta, tw1, tw2, tx, p);
  return cudaGetLastError();
}

}  // namespace fmlp

inline bool fused_mlp_supported(int C, int Hd) {
  return (C == 80 || C == 96 || C == 128 || C == 160 || C == 192 || C == 256 || C == 320) && Hd % 8 == 0 && Hd >= 64;
}
inline cudaError_t fused_mlp_configure_device() {
  cudaError_t e;
  if ((e = fmlp::configure_t<80>()) != cudaSuccess) return e;
  if ((e = fmlp::configure_t<96>()) != cudaSuccess) return e;
  if ((e = fmlp::configure_t<128>()) != cudaSuccess) return e;
  if ((e = fmlp::configure_t<160>()) != cudaSuccess) return e;
  if ((e = fmlp::configure_t<192>()) != cudaSuccess) return e;
  if ((e = fmlp::configure_t<256>()) != cudaSuccess) return e;
  return fmlp::configure_t<320>();
}
// a [M, C] bf16, w1 [Hd, C] bf16, w2 [C, Hd] bf16 (both torch Linear / 1x1-conv layout, K contiguous), x [M, C] fp32 in place
inline cudaError_t fused_mlp(const __nv_bfloat16* a, long long lda, const __nv_bfloat16* w1, long long ldw1, const float* b1,
                             const __nv_bfloat16* w2, long long ldw2, const float* b2, const float* gamma, float* x,
                             long long ldx, int M, int C, int Hd, int num_sms, cudaStream_t st) {
  if (M <= 0) return cudaSuccess;
  if (!fused_mlp_supported(C, Hd) || (lda & 7) || (ldw1 & 7) || (ldw2 & 7) || (ldx & 3)) return cudaErrorInvalidValue;
  fmlp::Params p;
  p.M = M; p.C = C; p.Hd = Hd; p.b1 = b1; p.b2 = b2; p.gamma = gamma; p.x = x; p.ldx = ldx;
#ifdef CLIPB200_FMLP_TIMING
  p.dbg = fmlp::timing_buffer();
#endif
#define CLIPB200_FMLP_CASE(C_) \
  if (C == C_) return fmlp::launch_t<C_>(a, lda, w1, ldw1, w2, ldw2, x, ldx, p, num_sms, st);
  CLIPB200_FMLP_CASE(80)
  CLIPB200_FMLP_CASE(96)
  CLIPB200_FMLP_CASE(128)
  CLIPB200_FMLP_CASE(160)
  CLIPB200_FMLP_CASE(192)
  CLIPB200_FMLP_CASE(256)
  CLIPB200_FMLP_CASE(320)
#undef CLIPB200_FMLP_CASE
  return cudaErrorInvalidValue;
}

}  // namespace clipb200
