// tcgen05 / TMEM flash attention (forward) for sm_100a.
//
//   out[b, q, h, :] = softmax(Q K^T / sqrt(hd)) V      qkv: [B*T, 3*H*hd] bf16 (q | k | v), out: [B*T, H*hd] bf16
//
// Replaces the MatMul -> Mul -> (+mask) -> Softmax -> MatMul chain ONNX Runtime executes inside `session.run`
// (reference src/vision.rs:108, src/text.rs:157-160).
//
// Persistent CTAs (2 per SM, 192 threads): warps 0..3 = softmax (one query row per thread), warp 4 = TMA producer,
// warp 5 = MMA issuer.  Work item = (batch, head, 128-query tile); K/V stream through a 2- or 3-stage smem ring in
// blocks of BKV keys.
//   S = Q K^T   : tcgen05.mma SS, M=128, N=BKV, accumulator S in TMEM (fp32)
//   softmax     : tcgen05.ld S -> registers, online max / exp2 / sum in fp32 (packed f32x2 arithmetic, a quarter of the
//                 exponentials as an FMA-pipe polynomial), P written back to TMEM as packed bf16
//   O += P V    : tcgen05.mma with the A operand (P) read from TMEM, V from smem as an MN-major operand
//   epilogue    : O / l -> bf16 -> smem -> TMA store
// Two protocols between the MMA warp and the softmax warps (template parameter DB, chosen per head dim by measurement):
// one S tile + separate P tile with s_full / s_empty / p_full / pv_done hand-offs, or two S tiles with P written in
// place (see Cfg).  Measured with every MMA and all softmax arithmetic removed (CLIPB200_ATTN_DBG=63), the kernel still
// takes 82 % of its run time: it is bound by the latency of the tcgen05.mma -> commit -> mbarrier -> tcgen05.ld round
// trips of each key block, not by MUFU, tensor or L2 throughput (profiles/r01g_attn_summary.md).
// Head dims that are not a multiple of 64 (72, 80, 96) are split into a 64-wide part (128-byte-swizzled tiles) and a
// remainder of 8-element chunk planes (no-swizzle "interleaved" tiles, zero plane appended when the remainder is not
// a multiple of 16), so neither the GEMMs nor HBM ever see padding.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <vector>

#include "gemm_sm100.cuh"  // get_encode_tiled
#include "ptx_sm100.cuh"

namespace clipb200 {

namespace attn {

#ifndef CLIPB200_ATTN_DBG
#define CLIPB200_ATTN_DBG 0  // profiling aid (tests/native/attn_test.cu): 1 no exp, 2 no S load, 4 no P store,
                             // 8 no remainder-plane MMAs, 16 no PV MMAs, 32 no QK MMAs (results wrong by construction)
#endif

#ifdef CLIPB200_ATTN_TIMING
// profiling aid: cycles each role of CTA 0 spends waiting on each barrier (read back by tests/native/attn_test.cu)
__device__ unsigned long long g_attn_wait[16];
#define ATTN_TIMED_WAIT(slot, bar, par)                                          \
  do {                                                                           \
    const long long t0_ = clock64();                                             \
    ptx::mbar_wait(bar, par);                                                    \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0)                              \
      atomicAdd(&g_attn_wait[slot], (unsigned long long)(clock64() - t0_));      \
  } while (0)
#else
#define ATTN_TIMED_WAIT(slot, bar, par) ptx::mbar_wait(bar, par)
#endif

constexpr int BQ = 128;
// SPLIT softmax warps share one query row (each takes BKV/SPLIT key columns): 4*SPLIT softmax warps + TMA + MMA warp.
// SPLIT = 2 doubles the independent instruction streams per SM sub-partition (two CTAs -> four softmax warps each), which
// is what hides the tcgen05.ld -> exp2 -> tcgen05.st latency chain; the pair exchanges its row maximum through smem.
#ifndef CLIPB200_ATTN_SPLIT
#define CLIPB200_ATTN_SPLIT 1
#endif
constexpr int SPLIT = CLIPB200_ATTN_SPLIT;
constexpr int NSW = 4 * SPLIT;            // softmax warps
constexpr int WARP_TMA = NSW;
constexpr int WARP_MMA = NSW + 1;
constexpr int THREADS = 32 * (NSW + 2);

// DB = double-buffered S: two S tiles in TMEM with P(j) written in place over S(j) (packed bf16 in its first BKV/2
// columns).  QK^T of block j+2 is issued right after PV of block j (same buffer, ordered by the tensor pipe's in-order
// execution), so the MMA warp never waits for the softmax warps to pick S up and the softmax warps never wait for a PV
// to retire before storing P: the two barrier round trips that serialised every block of the single-S protocol
// (s_full -> s_empty -> QK and p_full -> pv_done -> P store; with all arithmetic removed that skeleton alone took 82 % of
// the kernel's run time) overlap with the arithmetic instead.
// VT = V arrives TRANSPOSED ([B, H*hd, T], written by the qkv GEMM's epilogue): a V block is then a K-major operand
// (rows = head-dim columns of O, 128-byte rows of 64 keys + an optional 64-byte-row part of 32 keys) and every 16-key
// step of O += P V is ONE tcgen05.mma with N = HDP, instead of an N = 64 instruction plus an N = REMP instruction on
// MN-major tiles.  The tensor pipe's front end costs ~94 cycles per TMEM-operand MMA whatever N is
// (tests/native/mma_latency_test.cu), and those instructions were 72 of the 102 per 576-key tile.
template <int HD, int BKV, bool DB = false, bool VT = false>
struct Cfg {
  static_assert(HD >= 64 && HD % 8 == 0 && HD <= 128, "head dim");
  static_assert(BKV % 16 == 0 && BKV >= 32 && BKV <= 128, "kv block");
  static_assert(!VT || BKV % 64 == 0 || BKV % 64 == 32, "transposed V: 64-key atoms plus at most one 32-key part");
  static constexpr int REM = HD - 64;                    // elements beyond the 64-wide main part
  static constexpr int REMP = (REM + 15) / 16 * 16;      // padded to the MMA K granularity
  static constexpr int REM_PLANES = REM / 8;             // chunk planes TMA fills
  static constexpr int REMP_PLANES = REMP / 8;           // chunk planes the MMA reads
  static constexpr int HDP = 64 + REMP;
  static constexpr int Q_MAIN = BQ * 128;                // bytes
  static constexpr int Q_REM = REMP_PLANES * BQ * 16;
  static constexpr int KV_MAIN = BKV * 128;
  static constexpr int KV_REM = REMP_PLANES * BKV * 16;
  static constexpr int KV_TILE = KV_MAIN + KV_REM;       // K (and V in its natural layout)
  // transposed V block: V_ATOMS tiles of [HDP rows][64 keys] (128B swizzle) + one of [HDP rows][32 keys] (64B swizzle)
  static constexpr int V_ATOMS = BKV / 64;
  static constexpr int V_REM32 = BKV % 64;
  static constexpr int VT_ATOM = HDP * 128;
  static constexpr int VT_TILE = (V_ATOMS * VT_ATOM + (V_REM32 ? HDP * 64 : 0) + 1023) / 1024 * 1024;
  static constexpr int V_TILE = VT ? VT_TILE : KV_TILE;
  static constexpr int KV_STAGE = KV_TILE + V_TILE;
  static constexpr int K_TX = KV_MAIN + REM_PLANES * BKV * 16;   // bytes TMA delivers per K block
  static constexpr int V_TX = VT ? HD * BKV * 2 : K_TX;          // ... per V block
  static constexpr int Q_TX = Q_MAIN + REM_PLANES * BQ * 16;
  static constexpr int STAGES = (DB && HD < 96) ? 3 : 2;   // 3 where two CTAs still fit in one SM's shared memory
  static constexpr int OUT_ROW = HD * 2;                 // bytes
  static constexpr int OUT_WARP = 32 * OUT_ROW;
  // smem carve (offsets from a 1024-aligned base)
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_KV = (Q_MAIN + Q_REM + 1023) / 1024 * 1024;
  static constexpr int KV_STAGE_AL = (KV_STAGE + 1023) / 1024 * 1024;
  static constexpr int OFF_OUT = OFF_KV + STAGES * KV_STAGE_AL;
  static constexpr int OFF_XCH = OFF_OUT + (4 * OUT_WARP + 127) / 128 * 128;   // row-max / row-sum exchange
  static constexpr int XCH_BYTES = (2 * 2 + 2) * BQ * 4;
  static constexpr int OFF_BAR = OFF_XCH + XCH_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int CW = BKV / SPLIT;          // key columns per softmax thread
  static constexpr int OW = HDP / SPLIT;          // O columns per softmax thread (rescale / epilogue)
  static_assert(CW % 16 == 0 && OW % 8 == 0, "split granularity");
  // TMEM columns
  static constexpr int COL_S = 0;                 // DB: buffer i at COL_S + i * BKV
  static constexpr int COL_P = DB ? 0 : BKV;      // packed bf16: BKV/2 columns (DB: in place over the S buffer)
  static constexpr int COL_O = DB ? 2 * BKV : BKV + BKV / 2;
  static_assert(!DB || SPLIT == 1, "P aliases S: one softmax warp must own all columns of its rows");
  static constexpr int TMEM_COLS = 256;
  static_assert(COL_O + HDP <= TMEM_COLS, "TMEM budget (2 CTAs per SM -> 256 columns each)");
  static_assert(KV_MAIN % 1024 == 0, "swizzle atom alignment");
  static_assert(!VT || (KV_TILE % 1024 == 0 && VT_ATOM % 1024 == 0), "transposed V tiles start on swizzle-atom boundaries");
};

struct Params {
  int T, H, B;
  int q_tiles, n_items;
  float scale_log2e;
};

// ---- descriptors ------------------------------------------------------------------------------------------------
// no-swizzle ("interleaved") operand tile made of chunk planes [plane][row][8 elements]:
//   K-major  (Q, K):  LBO = plane stride (next 8 elements along K), SBO = 128 B (next 8 rows)
//   MN-major (V)   :  LBO = 128 B (next 8 keys along K),           SBO = plane stride (next 8 elements along N)
__device__ __forceinline__ uint64_t make_nosw_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // version
  return d;                             // layout type 0 = no swizzle
}
// MN-major, 128-byte swizzle: rows (keys) are 128 B = 64 elements along N, 8-row groups 1024 B apart (SBO);
// LBO (next 64-element block along N) is unused for N = 64.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// K-major, 64-byte swizzle: rows of 64 B = 32 elements along K, 8-row groups 512 B apart (SBO)
__device__ __forceinline__ uint64_t make_kmajor_sw64_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;   // SWIZZLE_64B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(b_mn_major) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const void* tmap, uint64_t* bar, void* smem_dst, int32_t c0, int32_t c1,
                                            int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(ptx::smem_u32(bar)), "r"(c0),
        "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tmap, const void* smem_src, int32_t c0, int32_t c1,
                                             int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(tmap)), "r"(ptx::smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x32_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :
               : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void pair_barrier(int quarter) {  // the SPLIT warps that share a row block
  asm volatile("bar.sync %0, %1;" ::"r"(quarter + 1), "n"(32 * SPLIT) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---- softmax arithmetic options (measured with tests/native/attn_test.cu; defaults = the fastest measured) -------------
// CLIPB200_ATTN_POLY_NUM / _DEN: of every DEN element pairs, NUM take exp2 on the FMA pipe (Cody-Waite split + degree-3
// minimax polynomial in packed f32x2 arithmetic: FFMA2 / FADD2 + one LEA per element, relative error 7.5e-5, far below
// the bf16 rounding of P) instead of MUFU.EX2.  The softmax warps are bound by the 16 exp2/clk/SM MUFU rate at head dim
// 72, so moving a fraction of the exponentials to the (otherwise ~half idle) FMA pipe shortens their critical path.
#ifndef CLIPB200_ATTN_POLY_NUM
#define CLIPB200_ATTN_POLY_NUM 1
#endif
#ifndef CLIPB200_ATTN_POLY_DEN
#define CLIPB200_ATTN_POLY_DEN 4
#endif
// CLIPB200_ATTN_PACKED: scale-and-shift (FFMA2) and row sums (FADD2) of the MUFU elements in packed f32x2 as well.
#ifndef CLIPB200_ATTN_PACKED
#define CLIPB200_ATTN_PACKED 1
#endif
// CLIPB200_ATTN_ELECT_ARRIVE: one lane per softmax warp arrives on s_empty / p_full (barrier count = warps) instead of
// all 32 lanes (count = threads): 31 fewer same-address shared-memory atomics per warp and barrier.
#ifndef CLIPB200_ATTN_ELECT_ARRIVE
#define CLIPB200_ATTN_ELECT_ARRIVE 0   // measured: no gain (634 -> 633 TFLOP/s at head dim 72)
#endif
// CLIPB200_ATTN_LDPIPE: the S tile is read from TMEM in 32-column chunks and the row maximum of chunk c is taken while
// chunk c+1 is in flight (TMEM reads run at 64 B/clk, so 96 columns take ~190 cycles that would otherwise be exposed).
#ifndef CLIPB200_ATTN_LDPIPE
#define CLIPB200_ATTN_LDPIPE 0         // measured: no gain (634 -> 631 TFLOP/s): the S read is not on the critical path
#endif
// CLIPB200_ATTN_MAXTREE: row maximum with four independent FMNMX3 chains instead of one 48-deep dependent chain.
#ifndef CLIPB200_ATTN_MAXTREE
#define CLIPB200_ATTN_MAXTREE 1
#endif
// CLIPB200_ATTN_PEXP=2 (experiment, off): where the head dim leaves a spare O column (hd 72 -> 80), P = exp2(s * scale - m)
// is produced by one packed `ex2.approx.ftz.bf16x2` per PAIR of scores (argument rounded to bf16 first) and the row sum l
// is not accumulated by the softmax warps at all: the first padding column of V (column HD of the zero plane) is set
// to 1.0, so O[:, HD] = sum_k P[:, k] comes out of the PV product in fp32.  Per score pair that is FFMA2 + CVT + 2 MUFU
// (the packed ex2 still issues one MUFU per half) against FFMA2 + 2 MUFU + FADD2 + F2FP or the 14-instruction polynomial
// pair: 35 % fewer instructions, but every exponential is back on the MUFU.  Measured on one box (profiles/r02j_*):
// correct on every case (max |error| vs fp32 0.0063 instead of 0.0038), 0.312 ms against 0.304 ms at T = 576 and 1.018
// against 0.985 ms at T = 2304 — slower: the MUFU rate, not the instruction count, is what the polynomial quarter buys
// back.  (The fp16 twin, P as fp16 with V in bf16, is rejected by the hardware: a tcgen05.mma.kind::f16 whose A and B
// formats differ raises "illegal instruction".)
#ifndef CLIPB200_ATTN_PEXP
#define CLIPB200_ATTN_PEXP 0
#endif
template <int HD, int BKV, bool DB, bool VT>
struct PExp {
  // needs the natural V layout's padding plane (columns HD .. HDP-1) and one softmax warp per row
  static constexpr bool kOn = CLIPB200_ATTN_PEXP != 0 && !VT && SPLIT == 1 && ((HD - 64) % 16) != 0 && HD > 64;
};
__device__ __forceinline__ uint32_t exp2_pair_bf16(float lo, float hi) {
  uint32_t h, r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(hi), "f"(lo));
  asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(r) : "r"(h));
  return r;
}
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// 2^x for a pair of arguments x <= ~9 without the MUFU: x = n + f with n = round(x), f in [-0.5, 0.5];
// 2^f by a degree-3 minimax polynomial, 2^n by adding n to the exponent field (the rounded integer sits in the low
// mantissa bits of x + 1.5 * 2^23).  Arguments are clamped at -125 so the exponent field cannot wrap: masked (-inf)
// scores come out as 2^-125 ~ 2e-38 instead of 0, which no bf16 / fp32 sum can see next to the row maximum's 2^0.
__device__ __forceinline__ void exp2_poly_pair(uint64_t x, float& p0, float& p1) {
  float x0, x1;
  unpack2(x, x0, x1);
  x = pack2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
  const uint64_t xf = add2(x, pack2(12582912.f, 12582912.f));
  const uint64_t xi = add2(xf, pack2(-12582912.f, -12582912.f));
  const uint64_t fr = fma2(xi, pack2(-1.f, -1.f), x);
  uint64_t p = fma2(pack2(0.0551716648f, 0.0551716648f), fr, pack2(0.2426111251f, 0.2426111251f));
  p = fma2(p, fr, pack2(0.6932609677f, 0.6932609677f));
  p = fma2(p, fr, pack2(0.9999280572f, 0.9999280572f));
  float q0, q1, f0, f1;
  unpack2(p, q0, q1);
  unpack2(xf, f0, f1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(f0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(f1) << 23));
}

// Per-thread view of one query tile's softmax pipeline (shared by the one-tile and the two-tile kernels).
struct SmxCtx {
  uint64_t *s_full, *s_empty, *p_full, *pv_done;
  uint32_t t_s, t_p, t_o;   // TMEM addresses of this thread's S / P / O slices (lane quarter and column slice applied)
  uint8_t* stg;             // this warp's [32][HD] bf16 output staging tile
  float *xch_max, *xch_sum; // SPLIT == 2 exchange buffers
  bool timed;               // this warp feeds the CLIPB200_ATTN_TIMING counters
};

// CLIPB200_ATTN_DEFER_EPI: the epilogue of a work item (O / l -> bf16 -> smem -> TMA store) is not run when the item's
// last P has been handed over, where the softmax warps would sit idle until the last PV retires (~1000 cycles of
// tcgen05.mma execution + commit + barrier hand-off), but inside the NEXT item's first key block, after that block's
// softmax arithmetic and right before its P store — the point where the protocol waits for that same PV anyway.  The next
// item's first PV (which overwrites O) is only issued once that P has been stored, i.e. after O has been read.
#ifndef CLIPB200_ATTN_DEFER_EPI
#define CLIPB200_ATTN_DEFER_EPI 1
#endif
// A/B switches for same-box measurements of the two other item-boundary changes (defaults = shipped behaviour):
// CLIPB200_ATTN_LOOKAHEAD=0: the first QK^T of an item is only issued after the previous item's last PV (single-S path);
// CLIPB200_ATTN_IDLE_SKIP=0: warps whose 32 query rows all lie beyond T run the full softmax on them anyway.
#ifndef CLIPB200_ATTN_LOOKAHEAD
#define CLIPB200_ATTN_LOOKAHEAD 1
#endif
#ifndef CLIPB200_ATTN_IDLE_SKIP
#define CLIPB200_ATTN_IDLE_SKIP 1
#endif
struct PendingEpilogue {
  bool valid = false;
  float l_run = 0.f;
  int qt = 0, h = 0, b = 0;
};

// O / l -> bf16 -> this warp's staging tile -> TMA store.  The caller has waited for the item's last PV.
template <int HD, int BKV, bool DB, bool VT>
__device__ __forceinline__ void epilogue_item(const SmxCtx& cx, const CUtensorMap* tm_out, float l_run, int qt, int h, int b,
                                              int quarter, int half, int lane) {
  using C = Cfg<HD, BKV, DB, VT>;
  constexpr int OW = C::OW;
  const int row = quarter * 32 + lane;
  ptx::tc_fence_after();
  if (SPLIT == 2) cx.xch_sum[half * BQ + row] = l_run;
  if (half == 0 && lane == 0) ptx::tma_store_wait_read<0>();  // previous item's store has left the staging tile
  if (SPLIT == 2) pair_barrier(quarter); else __syncwarp();
  float l_tot = SPLIT == 2 ? l_run + cx.xch_sum[(half ^ 1) * BQ + row] : l_run;
  if (PExp<HD, BKV, DB, VT>::kOn) {   // the row sum came out of the PV product: O[:, HD] (ones column of V)
    uint32_t r[8];
    tmem_ld_32x32_x8(cx.t_o + static_cast<uint32_t>(HD), r);
    ptx::tmem_ld_wait();
    l_tot = __uint_as_float(r[0]);
  }
  const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;
#pragma unroll
  for (int c = 0; c < OW / 8; ++c) {
    const int col = half * OW + c * 8;
    if (col < HD) {
      uint32_t r[8];
      tmem_ld_32x32_x8(cx.t_o + static_cast<uint32_t>(c * 8), r);
      ptx::tmem_ld_wait();
      uint4 o;
      o.x = pack_bf16(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv);
      o.y = pack_bf16(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv);
      o.z = pack_bf16(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv);
      o.w = pack_bf16(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv);
      *reinterpret_cast<uint4*>(cx.stg + lane * C::OUT_ROW + col * 2) = o;
    }
  }
  ptx::tc_fence_before();
  ptx::fence_proxy_async_smem();
  if (SPLIT == 2) pair_barrier(quarter); else __syncwarp();
  if (half == 0 && lane == 0) {
    tma_store_3d(tm_out, cx.stg, h * HD, qt * BQ + quarter * 32, b);
    ptx::tma_store_commit();
  }
}

// Softmax + epilogue of ONE work item (nb key/value blocks of one 128-query tile) for one softmax thread.
// `g` is the tile's running block counter (parity of s_full / s_empty / p_full / pv_done).
template <int HD, int BKV, bool CAUSAL, bool DB, bool VT>
__device__ __forceinline__ void softmax_item(const SmxCtx& cx, const Params& p, const CUtensorMap* tm_out, int qt, int h,
                                             int b, int nb, uint32_t& g, int quarter, int half, int lane,
                                             PendingEpilogue& pend) {
  using C = Cfg<HD, BKV, DB, VT>;
  constexpr int CW = C::CW, OW = C::OW;
  const int row = quarter * 32 + lane;
  const int qrow = qt * BQ + row;
  const bool idle_rows = CLIPB200_ATTN_IDLE_SKIP && SPLIT == 1 && qt * BQ + quarter * 32 >= p.T;   // warp-uniform
  float m_run = -INFINITY, l_run = 0.f;
  for (int j = 0; j < nb; ++j, ++g) {
    // DB: block g lives in S buffer g & 1; every per-buffer barrier completes one phase per two blocks
    const int buf = DB ? static_cast<int>(g & 1) : 0;
    const uint32_t par = DB ? ((g >> 1) & 1) : (g & 1);
    const uint32_t t_s = cx.t_s + static_cast<uint32_t>(buf * BKV);
    const uint32_t t_p = DB ? t_s : cx.t_p;
    if (cx.timed && lane == 0) { ATTN_TIMED_WAIT(8, &cx.s_full[buf], par); } else { ptx::mbar_wait(&cx.s_full[buf], par); }
    ptx::tc_fence_after();
    if (idle_rows) {
      // All 32 query rows of this warp lie beyond T (the lower half of the last tile when T % 128 <= 64, e.g. rows
      // 576..639 of a 576-token image): keep the hand-off protocol in step and do none of the arithmetic.  What the
      // PV then reads from this warp's P rows is whatever an earlier item left there; those O rows are never stored
      // (the output tensor map ends at T).  Frees a tenth of the softmax issue slots / MUFU time at T = 576.
      ptx::tc_fence_before();
      if (!DB) {
        if (CLIPB200_ATTN_ELECT_ARRIVE) { if (lane == 0) ptx::mbar_arrive(cx.s_empty); } else { ptx::mbar_arrive(cx.s_empty); }
      }
      uint64_t* prev_done_i = DB ? &cx.pv_done[(g - 1) & 1] : cx.pv_done;
      const uint32_t prev_par_i = DB ? (((g - 1) >> 1) & 1) : ((g & 1) ^ 1);
      const bool run_pending_i = CLIPB200_ATTN_DEFER_EPI && j == 0 && pend.valid;
      if (!DB || run_pending_i) {
        ptx::mbar_wait(prev_done_i, prev_par_i);
        ptx::tc_fence_after();
      }
      if (run_pending_i) {
        epilogue_item<HD, BKV, DB, VT>(cx, tm_out, pend.l_run, pend.qt, pend.h, pend.b, quarter, half, lane);
        pend.valid = false;
      }
      ptx::tc_fence_before();
      if (CLIPB200_ATTN_ELECT_ARRIVE) { if (lane == 0) ptx::mbar_arrive(&cx.p_full[buf]); } else { ptx::mbar_arrive(&cx.p_full[buf]); }
      continue;
    }
    float sv[CW];
    float mx_pipe[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // LDPIPE: maxima of the chunks already read
    if (CLIPB200_ATTN_LDPIPE && CW % 32 == 0 && !(CLIPB200_ATTN_DBG & 2)) {
      // chunk c+1 is requested before the maximum of chunk c is taken (tcgen05.wait::ld waits for every outstanding
      // load, so at most one chunk is in flight while the previous one is being reduced)
      uint32_t r32[CW / 32][32];
      ptx::tmem_ld_32x32(t_s, r32[0]);
#pragma unroll
      for (int c = 0; c < CW / 32; ++c) {
        ptx::tmem_ld_wait();
        if (c + 1 < CW / 32) ptx::tmem_ld_32x32(t_s + static_cast<uint32_t>((c + 1) * 32), r32[c + 1]);
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          sv[c * 32 + e] = __uint_as_float(r32[c][e]);
          mx_pipe[(e >> 1) & 3] = fmaxf(mx_pipe[(e >> 1) & 3], sv[c * 32 + e]);
        }
      }
    } else {
      // issue every TMEM load of this thread's slice before the single wait
      uint32_t r32[CW / 32 > 0 ? CW / 32 : 1][32];
      uint32_t r16[16];
#pragma unroll
      for (int c = 0; c < CW / 32; ++c) {
        if (CLIPB200_ATTN_DBG & 2) {
#pragma unroll
          for (int e = 0; e < 32; ++e) r32[c][e] = 0x3f800000u + e;
        } else {
          ptx::tmem_ld_32x32(t_s + static_cast<uint32_t>(c * 32), r32[c]);
        }
      }
      if (CW % 32 != 0) tmem_ld_32x32_x16(t_s + static_cast<uint32_t>(CW / 32 * 32), r16);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < CW / 32 * 32; ++e) sv[e] = __uint_as_float(r32[e / 32][e % 32]);
      if (CW % 32 != 0) {
#pragma unroll
        for (int e = 0; e < 16; ++e) sv[CW / 32 * 32 + e] = __uint_as_float(r16[e]);
      }
    }
    ptx::tc_fence_before();
    // S is in registers: the next QK^T may overwrite it
    if (!DB) {
      if (CLIPB200_ATTN_ELECT_ARRIVE) { if (lane == 0) ptx::mbar_arrive(cx.s_empty); } else { ptx::mbar_arrive(cx.s_empty); }
    }
    const int key0 = j * BKV + half * CW;
    const bool need_mask = (key0 + CW > p.T) || (CAUSAL && key0 + CW - 1 > qt * BQ + quarter * 32);
    if (need_mask) {  // warp-uniform, only the last kv block (and the diagonal blocks of causal towers)
#pragma unroll
      for (int e = 0; e < CW; ++e) {
        const int key = key0 + e;
        if (key >= p.T || (CAUSAL && key > qrow)) sv[e] = -INFINITY;
      }
    }
    // row maximum on the raw scores (the scale is positive, so max(s * scale) == scale * max(s))
    float mx = -INFINITY;
    if (CLIPB200_ATTN_LDPIPE && CW % 32 == 0 && !(CLIPB200_ATTN_DBG & 2) && !need_mask) {
      mx = fmaxf(fmaxf(mx_pipe[0], mx_pipe[1]), fmaxf(mx_pipe[2], mx_pipe[3]));
    } else if (CLIPB200_ATTN_MAXTREE) {
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int e = 0; e < CW; ++e) m4[(e >> 1) & 3] = fmaxf(m4[(e >> 1) & 3], sv[e]);
      mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    } else {
#pragma unroll
      for (int e = 0; e < CW; ++e) mx = fmaxf(mx, sv[e]);
    }
    if (SPLIT == 2) {  // combine with the warp that owns the other half of this row's columns
      float* slot = cx.xch_max + (g & 1) * 2 * BQ;
      slot[half * BQ + row] = mx;
      pair_barrier(quarter);
      mx = fmaxf(mx, slot[(half ^ 1) * BQ + row]);
    }
    mx *= p.scale_log2e;
    // Lazy rescaling: keep the running reference maximum unless the new block maximum exceeds it by more than
    // 2^8; P then stays <= 256 (exact in the fp32 sums, fine in bf16) and O / l are rescaled only rarely.
    // The result is mathematically identical because O and l always share the same reference maximum.
    const bool bump = (mx > m_run + 8.0f) || (m_run == -INFINITY);
    const float m_new = bump ? mx : m_run;
    const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
    const float alpha = bump ? ex2(m_run - m_use) : 1.0f;
    const float neg_m = -m_use;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pk[CW / 2];
    const uint64_t scale2 = pack2(p.scale_log2e, p.scale_log2e), negm2 = pack2(neg_m, neg_m);
    uint64_t rs2 = pack2(0.f, 0.f);
#pragma unroll
    for (int e = 0; e < CW / 2; ++e) {
      // exp2(s * scale - m): one FFMA + one MUFU per element, or (POLY_NUM of every POLY_DEN pairs) the packed
      // FMA-pipe polynomial
      float p0, p1;
      if (PExp<HD, BKV, DB, VT>::kOn) {
        float a0, a1;
        unpack2(fma2(pack2(sv[2 * e], sv[2 * e + 1]), scale2, negm2), a0, a1);
        pk[e] = exp2_pair_bf16(a0, a1);
        continue;
      }
      if (CLIPB200_ATTN_POLY_NUM > 0 && (e % CLIPB200_ATTN_POLY_DEN) >= CLIPB200_ATTN_POLY_DEN - CLIPB200_ATTN_POLY_NUM) {
        exp2_poly_pair(fma2(pack2(sv[2 * e], sv[2 * e + 1]), scale2, negm2), p0, p1);
      } else if (CLIPB200_ATTN_PACKED) {
        float a0, a1;
        unpack2(fma2(pack2(sv[2 * e], sv[2 * e + 1]), scale2, negm2), a0, a1);
        p0 = ex2(a0);
        p1 = ex2(a1);
      } else {
        const float a0 = fmaf(sv[2 * e], p.scale_log2e, neg_m), a1 = fmaf(sv[2 * e + 1], p.scale_log2e, neg_m);
        p0 = (CLIPB200_ATTN_DBG & 1) ? a0 * 0.001f : ex2(a0);
        p1 = (CLIPB200_ATTN_DBG & 1) ? a1 * 0.001f : ex2(a1);
      }
      if (CLIPB200_ATTN_PACKED) {
        rs2 = add2(rs2, pack2(p0, p1));
      } else {
        rs0 += p0;
        rs1 += p1;
      }
      pk[e] = pack_bf16(p0, p1);
    }
    if (CLIPB200_ATTN_PACKED) unpack2(rs2, rs0, rs1);
    l_run = l_run * alpha + (rs0 + rs1);  // partial over this warp's columns; the halves are added in the epilogue
    m_run = m_new;
    // P(j) and the O rescale must wait until PV(j-1) has finished reading P and writing O
    // DB: P(j) overwrites this thread's own (already loaded) S(j) columns, nothing to wait for
    const bool rescale = j > 0 && __any_sync(0xffffffffu, bump);
    // the previous block's PV: buffer (g - 1) & 1, phase (g - 1) >> 1
    uint64_t* prev_done = DB ? &cx.pv_done[(g - 1) & 1] : cx.pv_done;
    const uint32_t prev_par = DB ? (((g - 1) >> 1) & 1) : ((g & 1) ^ 1);
    const bool run_pending = CLIPB200_ATTN_DEFER_EPI && j == 0 && pend.valid;   // warp-uniform
    if (!DB || rescale || run_pending) {
      if (cx.timed && lane == 0) { ATTN_TIMED_WAIT(9, prev_done, prev_par); } else { ptx::mbar_wait(prev_done, prev_par); }
      ptx::tc_fence_after();
    }
    if (run_pending) {   // the previous item's last PV has retired: read its O before this item's first PV overwrites it
      epilogue_item<HD, BKV, DB, VT>(cx, tm_out, pend.l_run, pend.qt, pend.h, pend.b, quarter, half, lane);
      pend.valid = false;
    }
#pragma unroll
    for (int c = 0; c < CW / 32; ++c) {
      uint32_t r[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) r[e] = pk[c * 16 + e];
      if (CLIPB200_ATTN_DBG & 4) { if (r[0] == 0x12345678u && r[7] == 0x9abcdef0u) tmem_st_32x32_x16(t_p + static_cast<uint32_t>(c * 16), r); }
      else tmem_st_32x32_x16(t_p + static_cast<uint32_t>(c * 16), r);
    }
    if (CW % 32 != 0) {
      uint32_t r[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) r[e] = pk[CW / 32 * 16 + e];
      tmem_st_32x32_x8(t_p + static_cast<uint32_t>(CW / 32 * 16), r);
    }
    if (rescale) {  // rare: rescale this warp's slice of O
#pragma unroll
      for (int c = 0; c < OW / 16; ++c) {
        uint32_t r[16];
        tmem_ld_32x32_x16(cx.t_o + static_cast<uint32_t>(c * 16), r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
        tmem_st_32x32_x16(cx.t_o + static_cast<uint32_t>(c * 16), r);
      }
      if (OW % 16 != 0) {
        uint32_t r[8];
        tmem_ld_32x32_x8(cx.t_o + static_cast<uint32_t>(OW / 16 * 16), r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 8; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
        tmem_st_32x32_x8(cx.t_o + static_cast<uint32_t>(OW / 16 * 16), r);
      }
    }
    tmem_st_wait();
    ptx::tc_fence_before();
    if (CLIPB200_ATTN_ELECT_ARRIVE) { if (lane == 0) ptx::mbar_arrive(&cx.p_full[buf]); } else { ptx::mbar_arrive(&cx.p_full[buf]); }
  }
  if (idle_rows) {   // nothing of this warp's O rows is ever stored
    pend.valid = false;
    return;
  }
  if (CLIPB200_ATTN_DEFER_EPI) {   // handed to the next item's first block (or to the kernel's tail)
    pend.valid = true;
    pend.l_run = l_run;
    pend.qt = qt; pend.h = h; pend.b = b;
    return;
  }
  // epilogue: wait for the last PV, normalise, store this warp's slice of the O row
  {
    uint64_t* last_done = DB ? &cx.pv_done[(g - 1) & 1] : cx.pv_done;
    const uint32_t last_par = DB ? (((g - 1) >> 1) & 1) : ((g & 1) ^ 1);
    if (cx.timed && lane == 0) { ATTN_TIMED_WAIT(10, last_done, last_par); } else { ptx::mbar_wait(last_done, last_par); }
  }
  epilogue_item<HD, BKV, DB, VT>(cx, tm_out, l_run, qt, h, b, quarter, half, lane);
}

template <int HD, int BKV, bool CAUSAL, bool DB, bool VT>
__global__ void __launch_bounds__(THREADS, 2)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_q_main, const __grid_constant__ CUtensorMap tm_q_rem,
                        const __grid_constant__ CUtensorMap tm_kv_main, const __grid_constant__ CUtensorMap tm_kv_rem,
                        const __grid_constant__ CUtensorMap tm_vt_main, const __grid_constant__ CUtensorMap tm_vt_rem,
                        const __grid_constant__ CUtensorMap tm_out, Params p) {
  using C = Cfg<HD, BKV, DB, VT>;
  constexpr int ST = C::STAGES;
  extern __shared__ uint8_t attn_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(attn_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_q = smem + C::OFF_Q;
  uint8_t* s_kv = smem + C::OFF_KV;
  uint8_t* s_out = smem + C::OFF_OUT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  // K and V have separate full/empty barriers: a K slot is released as soon as its QK^T retires (long before the
  // matching PV), so the next K tile is prefetched about two iterations ahead even with a 2-deep ring.
  uint64_t* k_full = bars + 2;     // [ST <= 3]
  uint64_t* k_empty = bars + 5;    // [ST]
  uint64_t* v_full = bars + 8;     // [ST]
  uint64_t* v_empty = bars + 11;   // [ST]
  uint64_t* s_full = bars + 14;    // [2]  (single-S protocol: [0] only, likewise p_full / pv_done)
  uint64_t* p_full = bars + 16;    // [2]
  uint64_t* pv_done = bars + 18;   // [2]
  uint64_t* s_empty = bars + 20;   // single-S protocol only
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 21);
  static_assert(ST <= 3, "barrier carve");

  // warp index through a lane-0 broadcast: ptxas then treats it (and every role branch on it) as warp-uniform and keeps
  // the MMA warp's descriptors / TMEM addresses in uniform registers instead of moving each tcgen05.mma operand with R2UR
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int kv_blocks_total = (p.T + BKV - 1) / BKV;

  // zero the padding chunk planes (TMA never writes them) of Q and of both K/V stages
  if (C::REMP_PLANES > C::REM_PLANES) {
    for (int i = threadIdx.x; i < (C::REMP_PLANES - C::REM_PLANES) * BQ; i += THREADS)
      reinterpret_cast<uint4*>(s_q + C::Q_MAIN + C::REM_PLANES * BQ * 16)[i] = make_uint4(0, 0, 0, 0);
    for (int st = 0; st < C::STAGES; ++st)
      for (int kv = 0; kv < 2; ++kv) {
        uint8_t* tile = s_kv + st * C::KV_STAGE_AL + kv * C::KV_TILE;
        for (int i = threadIdx.x; i < (C::REMP_PLANES - C::REM_PLANES) * BKV; i += THREADS)
          reinterpret_cast<uint4*>(tile + C::KV_MAIN + C::REM_PLANES * BKV * 16)[i] = make_uint4(0, 0, 0, 0);
      }
    if (PExp<HD, BKV, DB, VT>::kOn) {
      __syncthreads();   // the zeroing above is complete
      for (int st = 0; st < C::STAGES; ++st) {
        uint8_t* vpad = s_kv + st * C::KV_STAGE_AL + C::KV_TILE + C::KV_MAIN + C::REM_PLANES * BKV * 16;
        for (int i = threadIdx.x; i < BKV; i += THREADS) *reinterpret_cast<uint16_t*>(vpad + i * 16) = 0x3F80;  // bf16 1.0
      }
    }
    ptx::fence_proxy_async_smem();
  }
  if (VT) {
    // rows HD..HDP-1 of the transposed V tiles (the O columns that pad the head dim to the MMA's N granularity) are
    // never written by TMA: zero the tiles once so the discarded accumulator columns stay finite
    for (int st = 0; st < C::STAGES; ++st) {
      uint4* tile = reinterpret_cast<uint4*>(s_kv + st * C::KV_STAGE_AL + C::KV_TILE);
      for (int i = threadIdx.x; i < C::V_TILE / 16; i += THREADS) tile[i] = make_uint4(0, 0, 0, 0);
    }
    ptx::fence_proxy_async_smem();
  }
  if (warp == WARP_TMA && lane == 0) {
    ptx::prefetch_tmap(&tm_q_main);
    ptx::prefetch_tmap(&tm_kv_main);
    ptx::prefetch_tmap(&tm_out);
    if (C::REM > 0) { ptx::prefetch_tmap(&tm_q_rem); ptx::prefetch_tmap(&tm_kv_rem); }
    if (VT) { ptx::prefetch_tmap(&tm_vt_main); if (C::V_REM32) ptx::prefetch_tmap(&tm_vt_rem); }
    ptx::mbar_init(q_full, 1);
    ptx::mbar_init(q_empty, 1);
    for (int s = 0; s < ST; ++s) {
      ptx::mbar_init(&k_full[s], 1); ptx::mbar_init(&k_empty[s], 1);
      ptx::mbar_init(&v_full[s], 1); ptx::mbar_init(&v_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&p_full[i], CLIPB200_ATTN_ELECT_ARRIVE ? NSW : 128 * SPLIT);
      ptx::mbar_init(&pv_done[i], 1);
    }
    ptx::mbar_init(s_empty, CLIPB200_ATTN_ELECT_ARRIVE ? NSW : 128 * SPLIT);
    ptx::fence_mbar_init();
  }
  if (warp == WARP_MMA) ptx::tmem_alloc<C::TMEM_COLS>(tmem_base_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_ptr, 0);   // warp-uniform (see `warp` above)

  auto item_blocks = [&](int qt) {
    if (!CAUSAL) return kv_blocks_total;
    const int last_q = qt * BQ + BQ - 1;
    const int nb = last_q / BKV + 1;
    return nb < kv_blocks_total ? nb : kv_blocks_total;
  };

  if (warp == WARP_TMA) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t g = 0;   // running K/V block counter
      uint32_t it = 0;  // running item counter
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        const int qt = item % p.q_tiles;
        const int bh = item / p.q_tiles;
        const int h = bh % p.H, b = bh / p.H;
        const int col_q = h * HD, col_k = p.H * HD + h * HD, col_v = 2 * p.H * HD + h * HD;
        ATTN_TIMED_WAIT(0, q_empty, (it & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(q_full, C::Q_TX);
        tma_load_3d(&tm_q_main, q_full, s_q, col_q, qt * BQ, b);
        for (int pl = 0; pl < C::REM_PLANES; ++pl)
          tma_load_3d(&tm_q_rem, q_full, s_q + C::Q_MAIN + pl * BQ * 16, col_q + 64 + 8 * pl, qt * BQ, b);
        const int nb = item_blocks(qt);
        for (int j = 0; j < nb; ++j, ++g) {
          const int st = g % ST;
          const uint32_t par = ((g / ST) & 1) ^ 1;
          uint8_t* kt = s_kv + st * C::KV_STAGE_AL;
          uint8_t* vt = kt + C::KV_TILE;
          ATTN_TIMED_WAIT(1, &k_empty[st], par);
          ptx::mbar_arrive_expect_tx(&k_full[st], C::K_TX);
          tma_load_3d(&tm_kv_main, &k_full[st], kt, col_k, j * BKV, b);
          for (int pl = 0; pl < C::REM_PLANES; ++pl)
            tma_load_3d(&tm_kv_rem, &k_full[st], kt + C::KV_MAIN + pl * BKV * 16, col_k + 64 + 8 * pl, j * BKV, b);
          ATTN_TIMED_WAIT(2, &v_empty[st], par);
          ptx::mbar_arrive_expect_tx(&v_full[st], C::V_TX);
          if (VT) {   // [B][H*hd][T]: rows h*HD .. +HD, keys j*BKV .. +BKV (keys beyond T are zero-filled)
            for (int a = 0; a < C::V_ATOMS; ++a)
              tma_load_3d(&tm_vt_main, &v_full[st], vt + a * C::VT_ATOM, j * BKV + a * 64, h * HD, b);
            if (C::V_REM32)
              tma_load_3d(&tm_vt_rem, &v_full[st], vt + C::V_ATOMS * C::VT_ATOM, j * BKV + C::V_ATOMS * 64, h * HD, b);
          } else {
            tma_load_3d(&tm_kv_main, &v_full[st], vt, col_v, j * BKV, b);
            for (int pl = 0; pl < C::REM_PLANES; ++pl)
              tma_load_3d(&tm_kv_rem, &v_full[st], vt + C::KV_MAIN + pl * BKV * 16, col_v + 64 + 8 * pl, j * BKV, b);
          }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, uniform control flow)
    // transposed V: descriptor of the 16-key step k of a block (k-steps 0..3 of each 64-key atom, then the 32-key part)
    constexpr uint32_t idesc_pv_vt = make_idesc(BQ, C::HDP, 0);
    auto vt_desc = [](uint32_t v_addr, int k) -> uint64_t {
      const int a = k >> 2;
      if (a < C::V_ATOMS) return ptx::make_kmajor_sw128_desc(v_addr + a * C::VT_ATOM) + static_cast<uint64_t>(2 * (k & 3));
      return make_kmajor_sw64_desc(v_addr + C::V_ATOMS * C::VT_ATOM) + static_cast<uint64_t>(2 * (k - 4 * C::V_ATOMS));
    };
    (void)idesc_pv_vt; (void)vt_desc;
    // One flat software pipeline over every key block this CTA will ever process: the QK^T of a block is issued a fixed
    // distance ahead of its PV (1 block with a single S tile, 2 with two), ACROSS work-item boundaries.  Q is single-
    // buffered, so the first QK^T of the next item waits for its Q tile, which the producer loads as soon as the current
    // item's last QK^T has retired (q_empty is committed right behind that last QK^T).  Without the look-ahead the first
    // S of every item was only requested after the previous item's last PV had been issued, and the softmax warps sat
    // idle for a PV + QK^T + commit round trip (~1100 cycles of a ~17 000-cycle item at T = 576).
    constexpr uint32_t idesc_qk = make_idesc(BQ, BKV, 0);
    constexpr uint32_t idesc_pv_main = make_idesc(BQ, 64, 1);
    constexpr uint32_t idesc_pv_rem = make_idesc(BQ, C::REMP > 0 ? C::REMP : 16, 1);
    (void)idesc_pv_main; (void)idesc_pv_rem;
    const uint32_t t_o = tmem_base + C::COL_O;
    const uint32_t q_addr = ptx::smem_u32(s_q);
    struct Cursor {   // (item, block-in-item) in processing order
      int item, j, nb;
      uint32_t it;
    };
    auto cursor_valid = [&](const Cursor& c) { return c.item < p.n_items; };
    auto cursor_begin = [&]() {
      Cursor c;
      c.item = blockIdx.x; c.j = 0; c.it = 0;
      c.nb = c.item < p.n_items ? item_blocks(c.item % p.q_tiles) : 0;
      return c;
    };
    auto cursor_next = [&](Cursor& c) {
      if (++c.j < c.nb) return;
      c.item += gridDim.x; c.j = 0; ++c.it;
      c.nb = c.item < p.n_items ? item_blocks(c.item % p.q_tiles) : 0;
    };
    Cursor cq = cursor_begin(), cp = cq;
    uint32_t gq = 0, gp = 0;   // running block counters of the QK^T and PV streams
    // S(gq) = Q K(gq)^T.  single S: waits until the softmax warps have read S(gq - 1); double S: no wait, the buffer's
    // previous contents (P of block gq - 2) were consumed by PV(gq - 2), issued earlier on the same in-order pipe.
    auto issue_next_qk = [&]() {
      if (!cursor_valid(cq)) return;
      if (cq.j == 0) {   // first block of an item: its Q tile must have landed
        ATTN_TIMED_WAIT(5, q_full, cq.it & 1);
      }
      const int st = DB ? static_cast<int>(gq % ST) : static_cast<int>(gq & 1);
      const int buf = DB ? static_cast<int>(gq & 1) : 0;
      const uint32_t k_addr = ptx::smem_u32(s_kv + st * C::KV_STAGE_AL);
      const uint32_t t_s = tmem_base + C::COL_S + static_cast<uint32_t>(buf * BKV);
      ATTN_TIMED_WAIT(3, &k_full[st], DB ? ((gq / ST) & 1) : ((gq >> 1) & 1));
      if (!DB) { ATTN_TIMED_WAIT(4, s_empty, (gq & 1) ^ 1); }
      ptx::tc_fence_after();
      const uint64_t dq = ptx::make_kmajor_sw128_desc(q_addr);
      const uint64_t dk = ptx::make_kmajor_sw128_desc(k_addr);
#pragma unroll
      for (int k = 0; k < ((CLIPB200_ATTN_DBG & 32) ? 0 : 4); ++k)
        ptx::umma_bf16_ss_w(t_s, dq + static_cast<uint64_t>(2 * k), dk + static_cast<uint64_t>(2 * k), idesc_qk,
                            k != 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < ((CLIPB200_ATTN_DBG & (8 | 32)) ? 0 : C::REMP / 16); ++k) {
        const uint64_t dqr = make_nosw_desc(q_addr + C::Q_MAIN + k * 2 * BQ * 16, BQ * 16, 128);
        const uint64_t dkr = make_nosw_desc(k_addr + C::KV_MAIN + k * 2 * BKV * 16, BKV * 16, 128);
        ptx::umma_bf16_ss_w(t_s, dqr, dkr, idesc_qk, 1u);
      }
      ptx::umma_commit_w(&k_empty[st]);
      ptx::umma_commit_w(&s_full[buf]);
      if (cq.j == cq.nb - 1) ptx::umma_commit_w(q_empty);   // every QK^T of this item is issued: Q is free when they retire
      ++gq;
      cursor_next(cq);
    };
    issue_next_qk();
    if (DB) issue_next_qk();
    while (cursor_valid(cp)) {
      // S(gp + 1) overlaps softmax(gp); without the look-ahead a block that starts a new item waits for this PV
      if (!DB && gq == gp + 1 && (CLIPB200_ATTN_LOOKAHEAD || cq.j != 0)) issue_next_qk();
      const int st = DB ? static_cast<int>(gp % ST) : static_cast<int>(gp & 1);
      const int buf = DB ? static_cast<int>(gp & 1) : 0;
      const uint32_t v_addr = ptx::smem_u32(s_kv + st * C::KV_STAGE_AL + C::KV_TILE);
      // P(gp): in place over S(gp) with two S tiles, in its own columns with one
      const uint32_t t_p = DB ? tmem_base + C::COL_S + static_cast<uint32_t>(buf * BKV) : tmem_base + C::COL_P;
      ATTN_TIMED_WAIT(6, &v_full[st], DB ? ((gp / ST) & 1) : ((gp >> 1) & 1));
      ATTN_TIMED_WAIT(7, &p_full[buf], DB ? ((gp >> 1) & 1) : (gp & 1));
      ptx::tc_fence_after();
#pragma unroll
      for (int k = 0; k < ((CLIPB200_ATTN_DBG & 16) ? 0 : BKV / 16); ++k) {
        const uint32_t acc = (cp.j | k) != 0 ? 1u : 0u;
        if (VT) {
          ptx::umma_bf16_ts_w(t_o, t_p + static_cast<uint32_t>(k * 8), vt_desc(v_addr, k), idesc_pv_vt, acc);
        } else {
          const uint64_t dv = make_mnmajor_sw128_desc(v_addr + k * 16 * 128);
          ptx::umma_bf16_ts_w(t_o, t_p + static_cast<uint32_t>(k * 8), dv, idesc_pv_main, acc);
          if (C::REMP > 0 && !(CLIPB200_ATTN_DBG & 8)) {
            const uint64_t dvr = make_nosw_desc(v_addr + C::KV_MAIN + k * 16 * 16, 128, BKV * 16);
            ptx::umma_bf16_ts_w(t_o + 64, t_p + static_cast<uint32_t>(k * 8), dvr, idesc_pv_rem, acc);
          }
        }
      }
      ptx::umma_commit_w(&v_empty[st]);
      ptx::umma_commit_w(&pv_done[buf]);
      ++gp;
      cursor_next(cp);
      if (DB) issue_next_qk();    // reuses the S buffer PV(gp - 1) has just been queued to consume
      else if (gq == gp) issue_next_qk();   // CLIPB200_ATTN_LOOKAHEAD == 0: first QK^T of the next item, behind the last PV
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue (warps 0..NSW-1)
    constexpr int CW = C::CW, OW = C::OW;
    const int quarter = warp & 3;     // TMEM lane quarter (query rows quarter*32 .. +31 of the tile)
    const int half = warp >> 2;       // which CW-wide slice of the key columns / OW-wide slice of O this warp owns
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    SmxCtx ctx;
    ctx.s_full = s_full; ctx.s_empty = s_empty; ctx.p_full = p_full; ctx.pv_done = pv_done;   // [2] each but s_empty
    ctx.t_s = tmem_base + lane_base + C::COL_S + static_cast<uint32_t>(half * CW);
    ctx.t_p = tmem_base + lane_base + C::COL_P + static_cast<uint32_t>(half * CW / 2);
    ctx.t_o = tmem_base + lane_base + C::COL_O + static_cast<uint32_t>(half * OW);
    ctx.stg = s_out + quarter * C::OUT_WARP;
    ctx.xch_max = reinterpret_cast<float*>(smem + C::OFF_XCH);  // [parity][half][row]
    ctx.xch_sum = ctx.xch_max + 2 * 2 * BQ;                     // [half][row]
    ctx.timed = warp == 0;
    uint32_t g = 0;
    PendingEpilogue pend;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int qt = item % p.q_tiles;
      const int bh = item / p.q_tiles;
      const int h = bh % p.H, b = bh / p.H;
      const int nb = item_blocks(qt);
      softmax_item<HD, BKV, CAUSAL, DB, VT>(ctx, p, &tm_out, qt, h, b, nb, g, quarter, half, lane, pend);
    }
    if (pend.valid) {   // the last item's epilogue has nobody to ride on
      uint64_t* last_done = DB ? &pv_done[(g - 1) & 1] : pv_done;
      const uint32_t last_par = DB ? (((g - 1) >> 1) & 1) : ((g & 1) ^ 1);
      ptx::mbar_wait(last_done, last_par);
      epilogue_item<HD, BKV, DB, VT>(ctx, &tm_out, pend.l_run, pend.qt, pend.h, pend.b, quarter, half, lane);
    }
    if (half == 0 && lane == 0) ptx::tma_store_wait<0>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

// 3-D tensor map over a [B, T, cols] bf16 tensor (cols contiguous), box = {box_cols, box_rows, 1}
inline bool make_tmap_3d(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t T, uint64_t B, uint32_t box_cols,
                         uint32_t box_rows, bool swizzle128) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return false;
  cuuint64_t dims[3] = {cols, T, B};
  cuuint64_t strides[2] = {cols * 2, cols * 2 * T};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}
// 3-D tensor map over the transposed V tensor [B, rows = H*hd, T keys] bf16 with a padded key stride `ld` (multiple
// of 8 elements): box = {box_keys, box_rows, 1}; 128-byte swizzle for 64-key boxes, 64-byte swizzle for 32-key boxes.
// The key dimension is T, not ld: keys beyond T are out of bounds (zero-filled on load, dropped on store).
inline bool make_tmap_vt(CUtensorMap* tm, const void* base, uint64_t T, uint64_t ld, uint64_t rows, uint64_t B,
                         uint32_t box_keys, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return false;
  cuuint64_t dims[3] = {T, rows, B};
  cuuint64_t strides[2] = {ld * 2, ld * 2 * rows};
  cuuint32_t box[3] = {box_keys, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_keys == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}
inline int attn_vt_ld(int T) { return (T + 7) & ~7; }   // key stride of the transposed V tensor

template <int HD, int BKV, bool CAUSAL, bool DB, bool VT>
inline cudaError_t launch_t(const __nv_bfloat16* qkv, const __nv_bfloat16* vt, __nv_bfloat16* out, int B, int T, int H,
                            int num_sms, cudaStream_t st) {
  using C = Cfg<HD, BKV, DB, VT>;
  // the seven tensor maps of a launch depend on (buffers, B, T, H) only: encoded once per host thread and instantiation
  struct Maps {
    const void *qkv, *vt, *out;
    int B, T, H;
    CUtensorMap q_main, q_rem, kv_main, kv_rem, vt_main, vt_rem, o_map;
  };
  static thread_local std::vector<Maps> cache;
  const Maps* m = nullptr;
  for (const Maps& c : cache)
    if (c.qkv == qkv && c.vt == vt && c.out == out && c.B == B && c.T == T && c.H == H) { m = &c; break; }
  if (m == nullptr) {
    Maps n;
    n.qkv = qkv; n.vt = vt; n.out = out; n.B = B; n.T = T; n.H = H;
    const uint64_t cols = 3ull * H * HD;
    if (!make_tmap_3d(&n.q_main, qkv, cols, T, B, 64, BQ, true)) return cudaErrorUnknown;
    if (!make_tmap_3d(&n.kv_main, qkv, cols, T, B, 64, BKV, true)) return cudaErrorUnknown;
    if (!make_tmap_3d(&n.q_rem, qkv, cols, T, B, 8, BQ, false)) return cudaErrorUnknown;
    if (!make_tmap_3d(&n.kv_rem, qkv, cols, T, B, 8, BKV, false)) return cudaErrorUnknown;
    n.vt_main = n.kv_main;
    n.vt_rem = n.kv_rem;
    if (VT) {
      if (vt == nullptr) return cudaErrorInvalidValue;
      if (!make_tmap_vt(&n.vt_main, vt, T, attn_vt_ld(T), static_cast<uint64_t>(H) * HD, B, 64, HD)) return cudaErrorUnknown;
      if (!make_tmap_vt(&n.vt_rem, vt, T, attn_vt_ld(T), static_cast<uint64_t>(H) * HD, B, 32, HD)) return cudaErrorUnknown;
    }
    if (!make_tmap_3d(&n.o_map, out, static_cast<uint64_t>(H) * HD, T, B, HD, 32, false)) return cudaErrorUnknown;
    if (cache.size() >= 64) cache.clear();
    cache.push_back(n);
    m = &cache.back();
  }
  const CUtensorMap &q_main = m->q_main, &q_rem = m->q_rem, &kv_main = m->kv_main, &kv_rem = m->kv_rem,
                    &vt_main = m->vt_main, &vt_rem = m->vt_rem, &o_map = m->o_map;
  Params p;
  p.T = T; p.H = H; p.B = B;
  p.q_tiles = (T + BQ - 1) / BQ;
  p.n_items = p.q_tiles * H * B;
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  const int grid = p.n_items < 2 * num_sms ? p.n_items : 2 * num_sms;
  attn_fwd_tcgen05_kernel<HD, BKV, CAUSAL, DB, VT><<<grid, THREADS, C::SMEM_BYTES, st>>>(q_main, q_rem, kv_main, kv_rem,
                                                                                         vt_main, vt_rem, o_map, p);
  return cudaGetLastError();
}
template <int HD, int BKV, bool DB, bool VT = false>
inline cudaError_t configure_t() {
  cudaError_t e = cudaFuncSetAttribute(attn_fwd_tcgen05_kernel<HD, BKV, false, DB, VT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<HD, BKV, DB, VT>::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(attn_fwd_tcgen05_kernel<HD, BKV, true, DB, VT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              Cfg<HD, BKV, DB, VT>::SMEM_BYTES);
}

// Double-buffered S needs one softmax warp per row block (P aliases S); the SPLIT == 2 experiment keeps the single-S
// protocol.
constexpr bool kDoubleS = SPLIT == 1;

}  // namespace attn

inline cudaError_t attn_tcgen05_configure_device() {
  cudaError_t e;
  if ((e = attn::configure_t<64, 96, false>()) != cudaSuccess) return e;
  if ((e = attn::configure_t<72, 96, false>()) != cudaSuccess) return e;
  if ((e = attn::configure_t<80, 96, false>()) != cudaSuccess) return e;
  if ((e = attn::configure_t<96, 64, false>()) != cudaSuccess) return e;
  // transposed-V variants (single S, 96-key blocks = one 64-key atom + one 32-key part per V block)
  if ((e = attn::configure_t<64, 96, false, true>()) != cudaSuccess) return e;
  if ((e = attn::configure_t<72, 96, false, true>()) != cudaSuccess) return e;
  if ((e = attn::configure_t<80, 96, false, true>()) != cudaSuccess) return e;
  if ((e = attn::configure_t<96, 64, false, true>()) != cudaSuccess) return e;
  if (attn::kDoubleS) {
    if ((e = attn::configure_t<64, 96, attn::kDoubleS, true>()) != cudaSuccess) return e;
    if ((e = attn::configure_t<72, 64, attn::kDoubleS, true>()) != cudaSuccess) return e;
    if ((e = attn::configure_t<80, 64, attn::kDoubleS, true>()) != cudaSuccess) return e;
    if ((e = attn::configure_t<96, 64, attn::kDoubleS, true>()) != cudaSuccess) return e;
  }
  if (attn::kDoubleS) {
    if ((e = attn::configure_t<64, 96, attn::kDoubleS>()) != cudaSuccess) return e;
    if ((e = attn::configure_t<72, 64, attn::kDoubleS>()) != cudaSuccess) return e;
    if ((e = attn::configure_t<80, 64, attn::kDoubleS>()) != cudaSuccess) return e;
    if ((e = attn::configure_t<96, 64, attn::kDoubleS>()) != cudaSuccess) return e;
  }
  return cudaSuccess;
}
inline bool attn_tcgen05_supported(int hd) { return hd == 64 || hd == 72 || hd == 80 || hd == 96; }
// Head dims / sequence lengths for which the transposed-V path (attn_tcgen05_vt + the qkv GEMM's EPI_QKVT epilogue) is
// the faster one (tests/native/attn_test.bin at T = 576, final kernels, profiles/r02d_attn_*.log):
//   hd 96: 755 vs 700 TFLOP/s   hd 72: 674 vs 676   hd 80: 735 vs 737   hd 64: 651 vs 673
// i.e. only where the natural layout needs two remainder planes per V block; the epilogue also needs T % 32 == 0.
inline bool attn_vt_preferred(int hd, int T) { return hd == 96 && T % 32 == 0; }

// qkv: [B*T, 3*H*hd] bf16, out: [B*T, H*hd] bf16
inline cudaError_t attn_tcgen05(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int H, int hd, bool causal,
                                int num_sms, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  // CLIPB200_ATTN_SINGLE_S=1 selects the single-S protocol (A/B measurements against the double-buffered default)
  static const bool single_s = !attn::kDoubleS || getenv("CLIPB200_ATTN_SINGLE_S") != nullptr;
#define CLIPB200_ATTN_CASE(HD_, BKV_, DB_)                                                             \
  if (hd == HD_)                                                                                       \
    return causal ? attn::launch_t<HD_, BKV_, true, DB_, false>(qkv, nullptr, out, B, T, H, num_sms, st) \
                  : attn::launch_t<HD_, BKV_, false, DB_, false>(qkv, nullptr, out, B, T, H, num_sms, st);
  // Which protocol per head dim is a measurement (tests/native/attn_test.cu, B = 128 / 64, T = 576, profiles/r01g_*):
  //   hd 64: double-buffered S, 96-key blocks   616 TFLOP/s  (single S: 589)
  //   hd 72: single S, 96-key blocks            634          (double S needs 64-key blocks to fit 256 TMEM columns: 617)
  //   hd 80: single S, 96-key blocks            678          (double S: 667)
  //   hd 96: double-buffered S, 64-key blocks   689          (single S: 680)
  if (!single_s) {
    CLIPB200_ATTN_CASE(64, 96, attn::kDoubleS)
    CLIPB200_ATTN_CASE(96, 64, attn::kDoubleS)
  }
  static const bool force_double_s = attn::kDoubleS && getenv("CLIPB200_ATTN_DOUBLE_S") != nullptr;
  if (force_double_s) {   // A/B switch for the other two head dims
    CLIPB200_ATTN_CASE(72, 64, attn::kDoubleS)
    CLIPB200_ATTN_CASE(80, 64, attn::kDoubleS)
  }
  // single-S kv block: 96 keys (576 = 6 x 96 exactly); 64 for head dim 96 so that two CTAs fit in one SM's shared memory
  CLIPB200_ATTN_CASE(64, 96, false)
  CLIPB200_ATTN_CASE(72, 96, false)
  CLIPB200_ATTN_CASE(80, 96, false)
  CLIPB200_ATTN_CASE(96, 64, false)
#undef CLIPB200_ATTN_CASE
  return cudaErrorInvalidValue;
}

// Same, with V taken from the transposed tensor vt [B, H*hd, ld = attn_vt_ld(T)] (written by the qkv GEMM's epilogue,
// gemm_sm100.cuh EPI_QKVT); the V third of `qkv` is not read.
inline cudaError_t attn_tcgen05_vt(const __nv_bfloat16* qkv, const __nv_bfloat16* vt, __nv_bfloat16* out, int B, int T,
                                   int H, int hd, bool causal, int num_sms, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
#define CLIPB200_ATTN_VT_CASE(HD_, BKV_, DB_)                                                                  \
  if (hd == HD_)                                                                                               \
    return causal ? attn::launch_t<HD_, BKV_, true, DB_, true>(qkv, vt, out, B, T, H, num_sms, st)             \
                  : attn::launch_t<HD_, BKV_, false, DB_, true>(qkv, vt, out, B, T, H, num_sms, st);
  // Protocol per head dim by measurement (tests/native/attn_test.bin, B = 128 / 64, T = 576, B200 at 1.965 GHz; r02c,
  // before the cross-item look-ahead): hd 72: single S 662 TFLOP/s (double S 648), hd 80: single S 713 (double S 703),
  // hd 96: double S 740 (single S 712), hd 64: 628 / 643.
  // CLIPB200_ATTN_DOUBLE_S=1 / CLIPB200_ATTN_SINGLE_S=1 force one protocol for every head dim (A/B runs).
  static const bool force_double = attn::kDoubleS && getenv("CLIPB200_ATTN_DOUBLE_S") != nullptr;
  static const bool force_single = !attn::kDoubleS || getenv("CLIPB200_ATTN_SINGLE_S") != nullptr;
  if (force_double) {
    CLIPB200_ATTN_VT_CASE(64, 96, attn::kDoubleS)
    CLIPB200_ATTN_VT_CASE(72, 64, attn::kDoubleS)
    CLIPB200_ATTN_VT_CASE(80, 64, attn::kDoubleS)
  }
  if (!force_single) {
    CLIPB200_ATTN_VT_CASE(96, 64, attn::kDoubleS)
  }
  CLIPB200_ATTN_VT_CASE(64, 96, false)
  CLIPB200_ATTN_VT_CASE(72, 96, false)
  CLIPB200_ATTN_VT_CASE(80, 96, false)
  CLIPB200_ATTN_VT_CASE(96, 64, false)
#undef CLIPB200_ATTN_VT_CASE
  return cudaErrorInvalidValue;
}

}  // namespace clipb200
