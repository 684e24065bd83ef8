// tcgen05 / TMEM / TMA GEMM for sm_100a:   C[M,N] = epilogue( A[M,K] . W[N,K]^T )
//
//  * A (activations) and W (torch Linear layout [out,in]) are both bf16, K contiguous ("K-major").
//  * Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread,
//    tcgen05.mma cta_group::1, UMMA 128 x BN x 16), warps 2..5 = epilogue (tcgen05.ld -> registers ->
//    fused bias / activation / residual / positional-embedding -> global).
//  * smem ring of STAGES x {A 128x64, W BNx64} bf16 tiles written by TMA with 128-byte swizzle.
//  * Two TMEM accumulators (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//  * M/N/K tails: TMA zero-fills out-of-bounds rows/columns; stores are guarded.  K%8==0, N%8==0.
//
// This replaces the MatMul/Gemm(+Add/+activation) nodes that ONNX Runtime executes inside
// `session.run` (reference src/vision.rs:108, src/text.rs:157-160).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <unordered_map>

#include "ptx_sm100.cuh"
#include "tensormap.h"

namespace clipb200 {

enum ActKind : int { ACT_NONE = 0, ACT_QUICKGELU = 1, ACT_GELU_TANH = 2, ACT_GELU_ERF = 3 };
enum EpiMode : int {
  EPI_BF16 = 0,   // out_bf16[r,c] = act(acc + bias[c])
  EPI_RESID = 1,  // out_f32[r,c] += gamma[c] * (acc + bias[c])          (fp32 residual stream, in place)
  EPI_F32 = 2,    // out_f32[remap(r),c] = acc + bias[c] + pos[pos_row(r),c]
  EPI_QKVT = 3    // EPI_BF16 for columns < vt_col0 (q | k); columns >= vt_col0 (v) are written TRANSPOSED into
                  // out_vt[b][c - vt_col0][t] with r = b * vt_T + t: the layout in which the attention kernel takes V
                  // as a K-major operand (one tcgen05.mma per 16 keys, attn_sm100.cuh).  Needs vt_T % 32 == 0.
};

struct GemmEpilogue {
  const float* bias = nullptr;   // [N] or null
  const float* gamma = nullptr;  // [N] or null (layer scale), EPI_RESID only
  const float* pos = nullptr;    // [rows_out, N] or null, EPI_F32 only
  __nv_bfloat16* out_bf16 = nullptr;
  float* out_f32 = nullptr;
  long long ldc = 0;  // elements
  int act = ACT_NONE;
  // EPI_F32 row remap: out_row = (r / rows_in) * rows_out + (r % rows_in) + row_off (rows_in == 0: identity)
  int rows_in = 0, rows_out = 0, row_off = 0;
  // EPI_QKVT: out_vt [vt_B][N - vt_col0][vt_ld] bf16, vt_T tokens per sequence (vt_ld >= vt_T, multiple of 8)
  __nv_bfloat16* out_vt = nullptr;
  int vt_col0 = 0, vt_T = 0, vt_ld = 0, vt_B = 0;
};

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ACT_QUICKGELU:
      return x / (1.0f + __expf(-1.702f * x));
    case ACT_GELU_TANH: {
      // 0.5 x (1 + tanh(u)) == x * sigmoid(2u),  u = sqrt(2/pi) (x + 0.044715 x^3)
      const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
      return x / (1.0f + __expf(-2.0f * u));
    }
    case ACT_GELU_ERF:
      return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f));
    default:
      return x;
  }
}

// tanh-form GELU on the MUFU.TANH path: 0.5 x (1 + tanh(u)); one special-function op per element.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Exact (erf) GELU without the erff() polynomial ladder: x * Phi(x) with Phi from Abramowitz-Stegun 7.1.26,
//   h = 0.5 * t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) * exp(-x^2 / 2),  t = 1 / (1 + p |x| / sqrt 2),
//   Phi(x) = x < 0 ? h : 1 - h          (|error of x * Phi| <= 4.3e-7 over [-12, 12], checked against scipy)
// = 8 FMA-pipe ops + MUFU.RCP + MUFU.EX2 per element instead of ~35: the FastViT fc1 epilogues were bound by it.
__device__ __forceinline__ float gelu_erf_fast(float x) {
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
__device__ __forceinline__ float apply_act_fast(float x, int act) {
  switch (act) {
    case ACT_QUICKGELU:
      return __fdividef(x, 1.0f + __expf(-1.702f * x));
    case ACT_GELU_TANH: {
      // u = sqrt(2/pi) (x + 0.044715 x^3) = x * (c1 + c2 x^2): 5 FMA-pipe ops + MUFU.TANH per element
      const float u = x * fmaf(x * x, 0.0356774081363001f, 0.7978845608028654f);
      const float hx = 0.5f * x;
      return fmaf(hx, tanh_approx(u), hx);
    }
    case ACT_GELU_ERF:
      return gelu_erf_fast(x);
    default:
      return x;
  }
}
__device__ __forceinline__ uint32_t pack2_act(float a, float b, int act) {
  if (act != ACT_NONE) {
    a = apply_act_fast(a, act);
    b = apply_act_fast(b, act);
  }
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

namespace ptx {
__device__ __forceinline__ void tmem_ld_32x32_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
}  // namespace ptx

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
// Epilogue warps per output mode (measured on B200, tests/native/gemm_test.bin and its mode 3):
//   EPI_BF16  -> 16 warps (4 per TMEM lane quarter), 32-column chunks, 64-byte staging rows with the 64B swizzle: half the
//                registers per thread and twice the independent instruction streams hide the tcgen05.ld -> bias /
//                activation -> st.shared -> TMA-store chain.  fc1 + GELU 1239 -> 1446 TFLOP/s, qkv 1370 -> 1454, the
//                short-K FastViT 1x1 convs +65 %.
//   EPI_RESID / EPI_F32 -> 8 warps (2 per quarter), 32-column chunks, 128-byte rows: with 16 warps the fp32 reduce-add
//                would need 16-column chunks (twice the TMA reduce operations) and measured 2-7 % slower.
// `CLIPB200_GEMM_EPI_WARPS` (8 | 16) overrides the bf16 choice for A/B runs.
#ifndef CLIPB200_GEMM_EPI_WARPS
#define CLIPB200_GEMM_EPI_WARPS 16
#endif
#ifndef CLIPB200_GEMM_RESID_WARPS
#define CLIPB200_GEMM_RESID_WARPS 8
#endif
template <int EPI>
struct EpiTraits {
  static constexpr bool BF16_OUT = EPI == 0 /* EPI_BF16 */ || EPI == 3 /* EPI_QKVT */;
  static constexpr int WARPS = BF16_OUT ? CLIPB200_GEMM_EPI_WARPS : (EPI == 1 ? CLIPB200_GEMM_RESID_WARPS : 8);
  static constexpr int SLOTS = WARPS / 4;                 // warps per TMEM lane quarter
  static constexpr bool NARROW = WARPS == 16 && BF16_OUT;
  static_assert(EPI != 3 || NARROW, "the transposing epilogue is written for the 16-warp bf16 layout");
  static constexpr int THREADS = 32 * (WARPS + 2);        // epilogue warps, then the TMA warp, then the MMA warp
  static constexpr int STAGE_BYTES = NARROW ? 32 * 64 : 32 * 128;  // per epilogue warp: 32 swizzled rows
};
// The two single-thread roles get the HIGHEST warp ids: the SM's warp arbiter favours higher warp ids, and a late
// tcgen05.mma / TMA issue starves the tensor pipe while the math-heavy epilogue warps can always wait a few cycles.
constexpr int GEMM_A_STAGE_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
constexpr int GEMM_EPI_STAGE_BYTES_F32MODE = 32 * 128;         // EPI_F32 transposes through a full 128-byte-row tile
// staging tiles per epilogue warp: with 2 the warp fills one tile while the TMA store / reduce of the previous chunk is
// still reading the other (cp.async.bulk.wait_group.read 1), at the price of one smem pipeline stage
#ifndef CLIPB200_GEMM_EPI_BUFS
#define CLIPB200_GEMM_EPI_BUFS 1
#endif
constexpr int GEMM_EPI_BUFS = CLIPB200_GEMM_EPI_BUFS;

// NCTA = 1: one CTA per 128 x BN tile.  NCTA = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x BN tile;
// each CTA stages its own 128 rows of A and HALF of the W tile, so the shared-memory traffic per MMA (TMA fill +
// operand read) drops from 1.5x to 1.0x of the 128 B/clk shared-memory bandwidth.
template <int BN, int NCTA = 1, int EPI = 0>
struct GemmCfg {
  static constexpr int B_STAGE_BYTES = (BN / NCTA) * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = GEMM_A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int EPI_WARP_BYTES = EPI == 2 /* EPI_F32 */ ? GEMM_EPI_STAGE_BYTES_F32MODE : GEMM_EPI_BUFS * EpiTraits<EPI>::STAGE_BYTES;
  static constexpr int EPI_BYTES = EpiTraits<EPI>::WARPS * EPI_WARP_BYTES;  // 32 KB in every mode
  static constexpr int BAR_BYTES = 256;
  static constexpr int BUDGET = 227 * 1024 - 1024 - BAR_BYTES - EPI_BYTES;
  static constexpr int STAGES = (BUDGET / STAGE_BYTES) > 6 ? 6 : (BUDGET / STAGE_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;  // +1024 alignment slack
};

template <int BN, int EPI, int NCTA>
__global__ void __launch_bounds__(EpiTraits<EPI>::THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                         const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_wt,
                         const __grid_constant__ CUtensorMap tmap_vt, int M, int N, int K, int n_tail, GemmEpilogue ep) {
  // n_tail = 128: the last column tile is only 128 wide (N = 1152 = 4 x 256 + 128, N = 3456 = 13 x 256 + 128): its W
  // box comes from `tmap_wt` and its MMAs use N = 128, so the tail costs half a tile instead of a padded full one.
  using Cfg = GemmCfg<BN, NCTA, EPI>;
  using ET = EpiTraits<EPI>;
  constexpr int GEMM_WARP_TMA = ET::WARPS, GEMM_WARP_MMA = ET::WARPS + 1;
  constexpr int TILE_M = GEMM_BM * NCTA;
  const int rank = NCTA == 2 ? static_cast<int>(ptx::cluster_ctarank()) : 0;      // CTA within the pair
  const int first_tile = NCTA == 2 ? static_cast<int>(ptx::cluster_id_x()) : static_cast<int>(blockIdx.x);
  const int tile_step = NCTA == 2 ? static_cast<int>(ptx::cluster_count_x()) : static_cast<int>(gridDim.x);
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * GEMM_A_STAGE_BYTES;
  uint8_t* smem_epi = smem + STAGES * Cfg::STAGE_BYTES;  // 1024-aligned (stage sizes are multiples of 1024)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + Cfg::EPI_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;           // [STAGES]  MMA -> TMA
  uint64_t* tmem_full_bar = bars + 2 * STAGES;   // [2]       MMA -> epilogue
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;  // [2]   epilogue -> MMA
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  // lane-0 broadcast: ptxas then knows the warp index (and every role branch on it) is warp-uniform, so the MMA warp's
  // descriptors and TMEM addresses live in uniform registers instead of being moved with R2UR before every tcgen05.mma
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + TILE_M - 1) / TILE_M;
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + GEMM_BK - 1) / GEMM_BK;

  if (warp == GEMM_WARP_TMA && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_w);
    if (EPI != EPI_F32) ptx::prefetch_tmap(&tmap_c);
    if (EPI == EPI_QKVT) ptx::prefetch_tmap(&tmap_vt);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], NCTA * ET::WARPS);  // one arrival per epilogue warp (of both CTAs)
    }
    ptx::fence_mbar_init();
  }
  if (warp == GEMM_WARP_MMA) {
    if (NCTA == 2) ptx::tmem_alloc_2sm<512>(tmem_base_ptr);
    else ptx::tmem_alloc<512>(tmem_base_ptr);
  }
  ptx::tc_fence_before();
  if (NCTA == 2) ptx::cluster_sync();  // the peer's barriers must be initialised before anything arrives remotely
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_ptr, 0);

  if (warp == GEMM_WARP_TMA) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          const bool tail_tile = n_tail != 0 && n_blk == n_tiles - 1;
          const int b_bytes = tail_tile ? (n_tail / NCTA) * GEMM_BK * 2 : Cfg::B_STAGE_BYTES;
          const CUtensorMap* tw = tail_tile ? &tmap_wt : &tmap_w;
          if (NCTA == 2) {
            // Both CTAs' loads are credited to the LEADER's full barrier; only the leader arrives (expecting the
            // bytes of both).  The peer never arrives: its bytes for the next phase can only land after the leader's
            // MMAs released the slot, and a transiently negative tx-count cannot complete a phase whose single
            // arrival is still pending.  (A remote release-arrive here costs a cluster-scope fence per k-block and
            // serialises the peer's TMA pipeline: measured 2x slower.)
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * (GEMM_A_STAGE_BYTES + b_bytes));
            ptx::tma_load_2d_2sm(&tmap_a, &full_bar[stage], smem_a + stage * GEMM_A_STAGE_BYTES, kb * GEMM_BK,
                                 m_blk * TILE_M + rank * GEMM_BM);
            ptx::tma_load_2d_2sm(tw, &full_bar[stage], smem_b + stage * Cfg::B_STAGE_BYTES, kb * GEMM_BK,
                                 n_blk * BN + rank * ((tail_tile ? n_tail : BN) / 2));
          } else {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], GEMM_A_STAGE_BYTES + b_bytes);
            ptx::tma_load_2d(&tmap_a, &full_bar[stage], smem_a + stage * GEMM_A_STAGE_BYTES, kb * GEMM_BK,
                             m_blk * GEMM_BM);
            ptx::tma_load_2d(tw, &full_bar[stage], smem_b + stage * Cfg::B_STAGE_BYTES, kb * GEMM_BK, n_blk * BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == GEMM_WARP_MMA) {
    // ------------------------------------------------------------ MMA issuer
    if (rank == 0) {  // in a pair only the leader CTA issues MMAs; whole warp in uniform control flow, one lane elected
                      // inside each issue (keeps descriptors in uniform registers: see ptx_sm100.cuh)
      constexpr uint32_t idesc_full = ptx::make_idesc_bf16_f32(TILE_M, BN);
      constexpr uint32_t idesc_tail = ptx::make_idesc_bf16_f32(TILE_M, 128);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t idesc = (n_tail != 0 && tile % n_tiles == n_tiles - 1) ? idesc_tail : idesc_full;
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * 256);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint64_t da = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem_a + stage * GEMM_A_STAGE_BYTES));
          const uint64_t db = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem_b + stage * Cfg::B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle atom: +2 in the (>>4) address field
            if (NCTA == 2)
              ptx::umma_bf16_ss_2sm_w(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                                    (kb | k) != 0 ? 1u : 0u);
            else
              ptx::umma_bf16_ss_w(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                                (kb | k) != 0 ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) when these MMAs retire
          if (NCTA == 2) ptx::umma_commit_2sm_w(&empty_bar[stage]);
          else ptx::umma_commit_w(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete (each CTA of a pair holds its own 128 rows in its own TMEM)
        if (NCTA == 2) ptx::umma_commit_2sm_w(&tmem_full_bar[acc]);
        else ptx::umma_commit_w(&tmem_full_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 0..7
    // Two warps per TMEM lane quarter, interleaved over column chunks.  Per chunk: tcgen05.ld (one accumulator row
    // per thread) -> bias / activation / layer-scale in registers -> 128B-swizzled smem staging tile (32 rows x 128 B)
    // -> one elected lane issues a TMA bulk tensor store (bf16 out) or a TMA reduce-add (fp32 residual stream), so
    // the global side is full-line, asynchronous and off the LSU.  The strided/remapped fp32 mode (patch embedding)
    // reads the staging tile back transposed and stores 4 rows x 128 B per warp instruction.
    const int ew = warp;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int half = ew >> 2;  // slot of this warp among the ET::SLOTS warps of its lane quarter
    uint8_t* const stg_base = smem_epi + ew * Cfg::EPI_WARP_BYTES;
    uint8_t* stg = stg_base;
    int stg_buf = 0;
    auto next_stg = [&]() {  // rotate to the staging tile whose store was issued longest ago
      if (GEMM_EPI_BUFS > 1) {
        stg_buf = (stg_buf + 1) % GEMM_EPI_BUFS;
        stg = stg_base + stg_buf * ET::STAGE_BYTES;
      }
    };
    const int sw = lane & 7;  // 128B-swizzle phase of this thread's staging row
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const int row_base = m_blk * TILE_M + rank * GEMM_BM + quarter * 32;
      const uint32_t taddr_row = tmem_base + static_cast<uint32_t>(acc * 256) +
                                 (static_cast<uint32_t>(quarter * 32) << 16);
      if ((EPI == EPI_BF16 || EPI == EPI_QKVT) && ET::NARROW) {
        // 16 epilogue warps: 32-column chunks, 64-byte staging rows (SWIZZLE_64B: 16-byte chunk ^= (row >> 1) & 3)
        const int sw64 = (lane >> 1) & 3;
#pragma unroll 1
        for (int c = half; c < BN / 32; c += ET::SLOTS) {
          const int n0 = n_blk * BN + c * 32;
          if (n0 >= N) break;  // warp-uniform
          uint32_t r[32];
          ptx::tmem_ld_32x32(taddr_row + static_cast<uint32_t>(c * 32), r);
          float4 b[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            b[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ep.bias != nullptr && n0 + 4 * j < N) b[j] = __ldg(reinterpret_cast<const float4*>(ep.bias + n0) + j);
          }
          ptx::tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            pk[2 * j] = pack2_act(__uint_as_float(r[4 * j]) + b[j].x, __uint_as_float(r[4 * j + 1]) + b[j].y, ep.act);
            pk[2 * j + 1] = pack2_act(__uint_as_float(r[4 * j + 2]) + b[j].z, __uint_as_float(r[4 * j + 3]) + b[j].w, ep.act);
          }
          if (lane == 0) ptx::tma_store_wait_read<0>();  // the previous store of this warp has left the staging tile
          __syncwarp();
          if (EPI == EPI_QKVT && n0 >= ep.vt_col0) {
            // V columns: the staging tile is written transposed, [32 columns][32 tokens] (64-byte rows, lane = token:
            // every 2-byte store instruction covers 64 contiguous bytes, conflict-free), and stored into
            // out_vt[b][column][t] by one TMA store.  (A token group that straddled two sequences would need a store
            // with a negative start coordinate for the second one; the hardware rejects that as an illegal
            // instruction — measured, r02b — so this mode requires T % 32 == 0.)
            uint16_t* st16 = reinterpret_cast<uint16_t*>(stg);
#pragma unroll
            for (int c2 = 0; c2 < 16; ++c2) {
              st16[(2 * c2) * 32 + lane] = static_cast<uint16_t>(pk[c2] & 0xFFFFu);
              st16[(2 * c2 + 1) * 32 + lane] = static_cast<uint16_t>(pk[c2] >> 16);
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              // vt_T is a multiple of 32 (checked by the launcher), so the 32 token rows of this warp belong to one
              // sequence; rows beyond M land at b >= vt_B and are dropped by the tensor map's bounds
              const int b = row_base / ep.vt_T;
              ptx::tma_store_3d(&tmap_vt, stg, row_base - b * ep.vt_T, n0 - ep.vt_col0, b);
              ptx::tma_store_commit();
            }
            continue;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((j ^ sw64) << 4)) =
                make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tmap_c, stg, n0, row_base);
            ptx::tma_store_commit();
          }
        }
      } else if (EPI == EPI_RESID && ET::NARROW) {
        const int sw64 = (lane >> 1) & 3;
#pragma unroll 1
        for (int c = half; c < BN / 16; c += ET::SLOTS) {
          const int n0 = n_blk * BN + c * 16;
          if (n0 >= N) break;  // warp-uniform
          uint32_t r[16];
          ptx::tmem_ld_32x32_x16(taddr_row + static_cast<uint32_t>(c * 16), r);
          float4 b4[4], g4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            b4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            g4[j] = make_float4(1.f, 1.f, 1.f, 1.f);
            if (n0 + 4 * j < N) {
              if (ep.bias != nullptr) b4[j] = __ldg(reinterpret_cast<const float4*>(ep.bias + n0) + j);
              if (ep.gamma != nullptr) g4[j] = __ldg(reinterpret_cast<const float4*>(ep.gamma + n0) + j);
            }
          }
          ptx::tmem_ld_wait();
          if (lane == 0) ptx::tma_store_wait_read<0>();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(stg + lane * 64 + ((j ^ sw64) << 4)) =
                make_float4((__uint_as_float(r[4 * j + 0]) + b4[j].x) * g4[j].x, (__uint_as_float(r[4 * j + 1]) + b4[j].y) * g4[j].y,
                            (__uint_as_float(r[4 * j + 2]) + b4[j].z) * g4[j].z, (__uint_as_float(r[4 * j + 3]) + b4[j].w) * g4[j].w);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_reduce_add_2d(&tmap_c, stg, n0, row_base);  // x[rows, cols] += tile, done in L2
            ptx::tma_store_commit();
          }
        }
      } else if (EPI == EPI_BF16) {
#pragma unroll 1
        for (int c = half; c < BN / 64; c += ET::SLOTS) {
          const int n0 = n_blk * BN + c * 64;
          if (n0 >= N) break;  // warp-uniform
          uint32_t r0[32], r1[32];
          ptx::tmem_ld_32x32(taddr_row + static_cast<uint32_t>(c * 64), r0);
          ptx::tmem_ld_32x32(taddr_row + static_cast<uint32_t>(c * 64 + 32), r1);
          float4 b0[8], b1[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            b0[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            b1[j] = b0[j];
            if (ep.bias != nullptr) {
              if (n0 + 4 * j < N) b0[j] = __ldg(reinterpret_cast<const float4*>(ep.bias + n0) + j);
              if (n0 + 32 + 4 * j < N) b1[j] = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + 32) + j);
            }
          }
          ptx::tmem_ld_wait();
          uint32_t pk[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            pk[2 * j] = pack2_act(__uint_as_float(r0[4 * j]) + b0[j].x, __uint_as_float(r0[4 * j + 1]) + b0[j].y, ep.act);
            pk[2 * j + 1] = pack2_act(__uint_as_float(r0[4 * j + 2]) + b0[j].z, __uint_as_float(r0[4 * j + 3]) + b0[j].w, ep.act);
            pk[16 + 2 * j] = pack2_act(__uint_as_float(r1[4 * j]) + b1[j].x, __uint_as_float(r1[4 * j + 1]) + b1[j].y, ep.act);
            pk[16 + 2 * j + 1] = pack2_act(__uint_as_float(r1[4 * j + 2]) + b1[j].z, __uint_as_float(r1[4 * j + 3]) + b1[j].w, ep.act);
          }
          next_stg();
          if (lane == 0) ptx::tma_store_wait_read<GEMM_EPI_BUFS - 1>();  // the store that used this tile has read it
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ sw) << 4)) =
                make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tmap_c, stg, n0, row_base);
            ptx::tma_store_commit();
          }
        }
      } else if (EPI == EPI_RESID) {
#pragma unroll 1
        for (int c = half; c < BN / 32; c += ET::SLOTS) {
          const int n0 = n_blk * BN + c * 32;
          if (n0 >= N) break;  // warp-uniform
          uint32_t r[32];
          ptx::tmem_ld_32x32(taddr_row + static_cast<uint32_t>(c * 32), r);
          float4 b4[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            b4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ep.bias != nullptr && n0 + 4 * j < N) b4[j] = __ldg(reinterpret_cast<const float4*>(ep.bias + n0) + j);
          }
          ptx::tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + b4[j].x;
            v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b4[j].y;
            v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b4[j].z;
            v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b4[j].w;
          }
          if (ep.gamma != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (n0 + 4 * j < N) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(ep.gamma + n0) + j);
                v[4 * j + 0] *= g.x; v[4 * j + 1] *= g.y; v[4 * j + 2] *= g.z; v[4 * j + 3] *= g.w;
              }
            }
          }
          next_stg();
          if (lane == 0) ptx::tma_store_wait_read<GEMM_EPI_BUFS - 1>();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ sw) << 4)) =
                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_reduce_add_2d(&tmap_c, stg, n0, row_base);  // x[rows, cols] += tile, done in L2
            ptx::tma_store_commit();
          }
        }
      } else {
        const int colq = lane & 7, rsub = lane >> 3;
#pragma unroll 1
        for (int c = half; c < BN / 32; c += ET::SLOTS) {
          const int n0 = n_blk * BN + c * 32;
          if (n0 >= N) break;  // warp-uniform
          uint32_t r[32];
          ptx::tmem_ld_32x32(taddr_row + static_cast<uint32_t>(c * 32), r);
          const int col = n0 + colq * 4;
          const bool col_ok = col < N;
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (col_ok && ep.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ sw) << 4)) =
                make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rl = 4 * i + rsub;
            const int grow = row_base + rl;
            float4 a = *reinterpret_cast<const float4*>(stg + rl * 128 + ((colq ^ (rl & 7)) << 4));
            a.x += b4.x; a.y += b4.y; a.z += b4.z; a.w += b4.w;
            if (ep.act != ACT_NONE) {  // 1x1 convolutions of the FastViT trunk that write the fp32 stream directly
              a.x = apply_act_fast(a.x, ep.act); a.y = apply_act_fast(a.y, ep.act);
              a.z = apply_act_fast(a.z, ep.act); a.w = apply_act_fast(a.w, ep.act);
            }
            if (grow < M && col_ok) {
              long long out_row = grow;
              int pos_row = 0;
              if (ep.rows_in > 0) {
                const int b = grow / ep.rows_in, t = grow - b * ep.rows_in;
                out_row = static_cast<long long>(b) * ep.rows_out + t + ep.row_off;
                pos_row = t + ep.row_off;
              }
              if (ep.pos != nullptr) {
                const float4 p4 = __ldg(reinterpret_cast<const float4*>(ep.pos + static_cast<long long>(pos_row) * N + col));
                a.x += p4.x; a.y += p4.y; a.z += p4.z; a.w += p4.w;
              }
              *reinterpret_cast<float4*>(ep.out_f32 + out_row * ep.ldc + col) = a;
            }
          }
          __syncwarp();
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // the accumulator buffer is owned by the (leader's) MMA thread
        if (rank == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
        else ptx::mbar_arrive_remote(&tmem_empty_bar[acc], 0);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (EPI != EPI_F32 && lane == 0) ptx::tma_store_wait<0>();  // all bulk stores of this warp are complete
  }

  ptx::tc_fence_before();
  if (NCTA == 2) ptx::cluster_sync();  // the leader's MMAs touch the peer's shared memory and TMEM until here
  else __syncthreads();
  if (warp == GEMM_WARP_MMA) {
    ptx::tc_fence_after();
    if (NCTA == 2) ptx::tmem_dealloc_2sm<512>(tmem_base);
    else ptx::tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
// Row-major [rows, cols] tensor (cols contiguous, leading dimension ld elements), box = [box_rows, 128 bytes of
// columns], 128-byte swizzle.  elem_bytes 2 = bf16 (64-column box), 4 = fp32 (32-column box).
inline bool make_tmap_2d_uncached(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                                  uint32_t box_rows, int elem_bytes, int inner_bytes);
// Encoding a tensor map is a driver call; the engine launches the same few hundred (pointer, shape) combinations every
// step (452 GEMM launches x 3-4 maps for one SO400M micro-batch), so the encoded maps are kept per host thread.
struct TmapKey {
  const void* base;
  uint64_t rows, cols, ld;
  uint32_t box_rows;
  int elem_bytes, inner_bytes;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows &&
           elem_bytes == o.elem_bytes && inner_bytes == o.inner_bytes;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = reinterpret_cast<uint64_t>(k.base) * 0x9E3779B97F4A7C15ull;
    h ^= (k.rows + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.cols * 0xBF58476D1CE4E5B9ull + (h << 6) + (h >> 2));
    h ^= (k.ld * 0x94D049BB133111EBull + (h << 6) + (h >> 2));
    h ^= ((static_cast<uint64_t>(k.box_rows) << 16 | static_cast<uint64_t>(k.elem_bytes) << 8 | static_cast<uint64_t>(k.inner_bytes)) + (h << 6) + (h >> 2));
    return static_cast<size_t>(h);
  }
};
inline bool make_tmap_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                         uint32_t box_rows, int elem_bytes, int inner_bytes = 128) {
  static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  const TmapKey key{base, rows, cols, ld, box_rows, elem_bytes, inner_bytes};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *tm = it->second;
    return true;
  }
  if (!make_tmap_2d_uncached(tm, base, rows, cols, ld, box_rows, elem_bytes, inner_bytes)) return false;
  if (cache.size() >= 4096) cache.clear();  // bounded: shapes change with the micro-batch tail, pointers with the engine
  cache.emplace(key, *tm);
  return true;
}
inline bool make_tmap_2d_uncached(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                                  uint32_t box_rows, int elem_bytes, int inner_bytes) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * static_cast<uint64_t>(elem_bytes)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(inner_bytes / elem_bytes), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   inner_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN, int EPI, int NCTA>
inline cudaError_t gemm_launch_t(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tc,
                                 const CUtensorMap& twt, const CUtensorMap& tvt, int n_tail, int M, int N, int K,
                                 const GemmEpilogue& ep, int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, NCTA, EPI>;
  const int tile_m = GEMM_BM * NCTA;
  const int tiles = ((M + tile_m - 1) / tile_m) * ((N + BN - 1) / BN);
  const int slots = num_sms / NCTA;
  const int groups = tiles < slots ? tiles : slots;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * NCTA);
  cfg.blockDim = dim3(EpiTraits<EPI>::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<BN, EPI, NCTA>, ta, tw, tc, twt, tvt, M, N, K, n_tail, ep);
}

template <int BN, int EPI, int NCTA>
inline cudaError_t gemm_configure_t() {
  return cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, EPI, NCTA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              GemmCfg<BN, NCTA, EPI>::SMEM_BYTES);
}

// Opt every instantiation into >48 KB dynamic shared memory on the CURRENT device (call once per device).
inline cudaError_t gemm_configure_device() {
  cudaError_t e;
#define CLIPB200_CFG(BN_)                                                      \
  if ((e = gemm_configure_t<BN_, EPI_BF16, 1>()) != cudaSuccess) return e;     \
  if ((e = gemm_configure_t<BN_, EPI_RESID, 1>()) != cudaSuccess) return e;    \
  if ((e = gemm_configure_t<BN_, EPI_F32, 1>()) != cudaSuccess) return e;      \
  if ((e = gemm_configure_t<BN_, EPI_QKVT, 1>()) != cudaSuccess) return e;     \
  if ((e = gemm_configure_t<BN_, EPI_QKVT, 2>()) != cudaSuccess) return e;     \
  if ((e = gemm_configure_t<BN_, EPI_BF16, 2>()) != cudaSuccess) return e;     \
  if ((e = gemm_configure_t<BN_, EPI_RESID, 2>()) != cudaSuccess) return e;    \
  if ((e = gemm_configure_t<BN_, EPI_F32, 2>()) != cudaSuccess) return e;
  CLIPB200_CFG(256)
  CLIPB200_CFG(192)
  CLIPB200_CFG(128)
#undef CLIPB200_CFG
  return cudaSuccess;
}

// Width of the last column tile when it is narrower than BN (only BN = 256 supports it): 128 if the remainder of N fits.
inline bool gemm_tail_enabled() {  // CLIPB200_GEMM_NO_TAIL=1 pads the last tile instead (A/B runs on one box)
  static const bool on = !(getenv("CLIPB200_GEMM_NO_TAIL") != nullptr && atoi(getenv("CLIPB200_GEMM_NO_TAIL")) != 0);
  return on;
}
inline int gemm_tail_cols(int N, int bn) {
  return (gemm_tail_enabled() && bn == 256 && N % 256 != 0 && N % 256 <= 128) ? 128 : 0;
}

inline int gemm_pick_bn(int N, int K) {
  // minimise (padded columns) x (relative cost per column of that tile width).  Measured on B200 with CTA pairs
  // (tests/native/gemm_test.bin 5, M = 147456): the 256-wide tile re-reads A less often and needs less shared-memory
  // bandwidth per MMA; per column the 192-wide tile costs 1.10x at K = 4304 and 1.15x at K = 1152 (so N = 1152 takes
  // 256-wide tiles with 10 % padding when K is short, 192-wide ones when K is long), the 128-wide tile 1.25x.
  const int cands[3] = {256, 192, 128};
  const double factor[3] = {1.0, K <= 2048 ? 1.15 : 1.10, 1.25};
  int best = 256;
  double best_cost = -1.0;
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    double cost = static_cast<double>((N + bn - 1) / bn) * bn * factor[i];
    // 256-wide tiles with a 128-wide tail tile (gemm_tail_cols): the tail costs 128 columns at the narrow-tile rate
    if (gemm_tail_cols(N, bn) != 0) cost = static_cast<double>(N / 256) * 256 + 128 * factor[2];
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

// A: [M,K] bf16 (lda elements), W: [N,K] bf16 (ldw elements).  force_bn: 0 = auto.
// force_ncta: 0 = auto (CTA pairs for large M), 1 / 2 = forced.
inline cudaError_t gemm_bf16(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw, int M,
                             int N, int K, int epi_mode, const GemmEpilogue& ep, int num_sms, cudaStream_t stream,
                             int force_bn = 0, int force_ncta = 0) {
  if (M <= 0 || N <= 0 || K <= 0) return cudaErrorInvalidValue;
  if ((K & 7) || (N & 7) || (lda & 7) || (ldw & 7) || (ep.ldc & 7)) return cudaErrorInvalidValue;
  const int bn = force_bn ? force_bn : gemm_pick_bn(N, K);
  CUtensorMap ta, tw, tc;
  if (!make_tmap_2d(&ta, A, M, K, lda, GEMM_BM, 2)) return cudaErrorUnknown;
  const int ncta = force_ncta ? force_ncta : (M >= 2048 ? 2 : 1);
  if (!make_tmap_2d(&tw, W, N, K, ldw, bn / ncta, 2)) return cudaErrorUnknown;
  const int n_tail = gemm_tail_cols(N, bn);
  CUtensorMap twt = tw;
  if (n_tail != 0 && !make_tmap_2d(&twt, W, N, K, ldw, n_tail / ncta, 2)) return cudaErrorUnknown;
  CUtensorMap tvt = ta;  // only read in EPI_QKVT mode
  if (epi_mode == EPI_BF16 || epi_mode == EPI_QKVT) {
    if (!make_tmap_2d(&tc, ep.out_bf16, M, N, ep.ldc, 32, 2, EpiTraits<EPI_BF16>::NARROW ? 64 : 128)) return cudaErrorUnknown;
    if (epi_mode == EPI_QKVT) {
      // columns [vt_col0, N) go to out_vt [B][N - vt_col0][ld] in 32-column x 32-token boxes (64-byte inner rows)
      if (ep.out_vt == nullptr || ep.vt_T <= 0 || (ep.vt_T & 31) || ep.vt_B <= 0 || (ep.vt_ld & 7) || ep.vt_ld < ep.vt_T ||
          (ep.vt_col0 & 31) || ep.vt_col0 >= N)
        return cudaErrorInvalidValue;
      PFN_encodeTiled enc = get_encode_tiled();
      if (enc == nullptr) return cudaErrorUnknown;
      const cuuint64_t rows = static_cast<cuuint64_t>(N - ep.vt_col0);
      cuuint64_t dims[3] = {static_cast<cuuint64_t>(ep.vt_T), rows, static_cast<cuuint64_t>(ep.vt_B)};
      cuuint64_t strides[2] = {static_cast<cuuint64_t>(ep.vt_ld) * 2, static_cast<cuuint64_t>(ep.vt_ld) * 2 * rows};
      cuuint32_t box[3] = {32, 32, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      if (enc(&tvt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ep.out_vt, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorUnknown;
    }
  } else if (epi_mode == EPI_RESID) {
    if (!make_tmap_2d(&tc, ep.out_f32, M, N, ep.ldc, 32, 4, EpiTraits<EPI_RESID>::NARROW ? 64 : 128)) return cudaErrorUnknown;
  } else {
    tc = ta;  // unused by the kernel in this mode
  }
#define CLIPB200_GEMM_CASE(BN_)                                                                                  \
  if (bn == BN_ && ncta == 1) {                                                                                  \
    if (epi_mode == EPI_BF16) return gemm_launch_t<BN_, EPI_BF16, 1>(ta, tw, tc, twt, tvt, n_tail, M, N, K, ep, num_sms, stream);   \
    if (epi_mode == EPI_QKVT) return gemm_launch_t<BN_, EPI_QKVT, 1>(ta, tw, tc, twt, tvt, n_tail, M, N, K, ep, num_sms, stream);   \
    if (epi_mode == EPI_RESID) return gemm_launch_t<BN_, EPI_RESID, 1>(ta, tw, tc, twt, tvt, n_tail, M, N, K, ep, num_sms, stream); \
    return gemm_launch_t<BN_, EPI_F32, 1>(ta, tw, tc, twt, tvt, n_tail, M, N, K, ep, num_sms, stream);                              \
  }                                                                                                              \
  if (bn == BN_ && ncta == 2) {                                                                                  \
    if (epi_mode == EPI_BF16) return gemm_launch_t<BN_, EPI_BF16, 2>(ta, tw, tc, twt, tvt, n_tail, M, N, K, ep, num_sms, stream);   \
    if (epi_mode == EPI_QKVT) return gemm_launch_t<BN_, EPI_QKVT, 2>(ta, tw, tc, twt, tvt, n_tail, M, N, K, ep, num_sms, stream);   \
    if (epi_mode == EPI_RESID) return gemm_launch_t<BN_, EPI_RESID, 2>(ta, tw, tc, twt, tvt, n_tail, M, N, K, ep, num_sms, stream); \
    return gemm_launch_t<BN_, EPI_F32, 2>(ta, tw, tc, twt, tvt, n_tail, M, N, K, ep, num_sms, stream);                              \
  }
  CLIPB200_GEMM_CASE(256)
  CLIPB200_GEMM_CASE(192)
  CLIPB200_GEMM_CASE(128)
#undef CLIPB200_GEMM_CASE
  return cudaErrorInvalidValue;
}

}  // namespace clipb200
