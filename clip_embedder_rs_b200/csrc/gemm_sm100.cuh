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

#include "ptx_sm100.cuh"

namespace clipb200 {

enum ActKind : int { ACT_NONE = 0, ACT_QUICKGELU = 1, ACT_GELU_TANH = 2, ACT_GELU_ERF = 3 };
enum EpiMode : int {
  EPI_BF16 = 0,   // out_bf16[r,c] = act(acc + bias[c])
  EPI_RESID = 1,  // out_f32[r,c] += gamma[c] * (acc + bias[c])          (fp32 residual stream, in place)
  EPI_F32 = 2     // out_f32[remap(r),c] = acc + bias[c] + pos[pos_row(r),c]
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

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_A_STAGE_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB

template <int BN>
struct GemmCfg {
  static constexpr int B_STAGE_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = GEMM_A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (200 * 1024 / STAGE_BYTES) > 6 ? 6 : (200 * 1024 / STAGE_BYTES);
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024 alignment slack
};

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                         int M, int N, int K, GemmEpilogue ep) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * GEMM_A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;           // [STAGES]  MMA -> TMA
  uint64_t* tmem_full_bar = bars + 2 * STAGES;   // [2]       MMA -> epilogue
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;  // [2]   epilogue -> MMA
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_w);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_base_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          ptx::tma_load_2d(&tmap_a, &full_bar[stage], smem_a + stage * GEMM_A_STAGE_BYTES, kb * GEMM_BK,
                           m_blk * GEMM_BM);
          ptx::tma_load_2d(&tmap_w, &full_bar[stage], smem_b + stage * Cfg::B_STAGE_BYTES, kb * GEMM_BK,
                           n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(GEMM_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * 256);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint64_t da = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem_a + stage * GEMM_A_STAGE_BYTES));
          const uint64_t db = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem_b + stage * Cfg::B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle atom: +2 in the (>>4) address field
            ptx::umma_bf16_ss(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                              (kb | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tmem_full_bar[acc]);  // accumulator complete
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 2..5
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const int row = m_blk * GEMM_BM + quarter * 32 + lane;
      const bool row_ok = row < M;
      long long out_row = row;
      int pos_row = 0;
      if (EPI == EPI_F32 && ep.rows_in > 0) {
        const int b = row / ep.rows_in, t = row - b * ep.rows_in;
        out_row = static_cast<long long>(b) * ep.rows_out + t + ep.row_off;
        pos_row = t + ep.row_off;
      }
      const uint32_t taddr_row = tmem_base + static_cast<uint32_t>(acc * 256) +
                                 (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int n0 = n_blk * BN + c * 32;
        if (n0 >= N) break;  // warp-uniform
        uint32_t r[32];
        ptx::tmem_ld_32x32(taddr_row + static_cast<uint32_t>(c * 32), r);
        ptx::tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (ep.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (n0 + 4 * j < N) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n0) + j);
              v[4 * j + 0] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
            }
          }
        }
        if (EPI == EPI_BF16) {
          if (ep.act != ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], ep.act);
          }
          if (row_ok) {
            __nv_bfloat16* dst = ep.out_bf16 + static_cast<long long>(row) * ep.ldc + n0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (n0 + 8 * j < N) {
                uint4 pk;
                __nv_bfloat162 h;
                h = __floats2bfloat162_rn(v[8 * j + 0], v[8 * j + 1]); pk.x = *reinterpret_cast<uint32_t*>(&h);
                h = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]); pk.y = *reinterpret_cast<uint32_t*>(&h);
                h = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]); pk.z = *reinterpret_cast<uint32_t*>(&h);
                h = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]); pk.w = *reinterpret_cast<uint32_t*>(&h);
                *reinterpret_cast<uint4*>(dst + 8 * j) = pk;
              }
            }
          }
        } else if (EPI == EPI_RESID) {
          if (row_ok) {
            float* dst = ep.out_f32 + static_cast<long long>(row) * ep.ldc + n0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (n0 + 4 * j < N) {
                float4 x = *reinterpret_cast<float4*>(dst + 4 * j);
                if (ep.gamma != nullptr) {
                  const float4 g = __ldg(reinterpret_cast<const float4*>(ep.gamma + n0) + j);
                  x.x += g.x * v[4 * j + 0]; x.y += g.y * v[4 * j + 1];
                  x.z += g.z * v[4 * j + 2]; x.w += g.w * v[4 * j + 3];
                } else {
                  x.x += v[4 * j + 0]; x.y += v[4 * j + 1]; x.z += v[4 * j + 2]; x.w += v[4 * j + 3];
                }
                *reinterpret_cast<float4*>(dst + 4 * j) = x;
              }
            }
          }
        } else {  // EPI_F32
          if (row_ok) {
            float* dst = ep.out_f32 + out_row * ep.ldc + n0;
            const float* pp = ep.pos != nullptr ? ep.pos + static_cast<long long>(pos_row) * N + n0 : nullptr;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (n0 + 4 * j < N) {
                float4 x = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                if (pp != nullptr) {
                  const float4 p4 = __ldg(reinterpret_cast<const float4*>(pp) + j);
                  x.x += p4.x; x.y += p4.y; x.z += p4.z; x.w += p4.w;
                }
                *reinterpret_cast<float4*>(dst + 4 * j) = x;
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tmem_empty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// bf16 row-major [rows, cols] (cols contiguous, leading dimension ld elements), box = [box_rows, 64 cols], SW128.
inline bool make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                              uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {GEMM_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN, int EPI>
inline cudaError_t gemm_launch_t(const CUtensorMap& ta, const CUtensorMap& tw, int M, int N, int K,
                                 const GemmEpilogue& ep, int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  const int tiles = ((M + GEMM_BM - 1) / GEMM_BM) * ((N + BN - 1) / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  gemm_bf16_tcgen05_kernel<BN, EPI><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tw, M, N, K, ep);
  return cudaGetLastError();
}

template <int BN, int EPI>
inline cudaError_t gemm_configure_t() {
  return cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              GemmCfg<BN>::SMEM_BYTES);
}

// Opt every instantiation into >48 KB dynamic shared memory on the CURRENT device (call once per device).
inline cudaError_t gemm_configure_device() {
  cudaError_t e;
#define CLIPB200_CFG(BN_)                                                   \
  if ((e = gemm_configure_t<BN_, EPI_BF16>()) != cudaSuccess) return e;     \
  if ((e = gemm_configure_t<BN_, EPI_RESID>()) != cudaSuccess) return e;    \
  if ((e = gemm_configure_t<BN_, EPI_F32>()) != cudaSuccess) return e;
  CLIPB200_CFG(256)
  CLIPB200_CFG(192)
  CLIPB200_CFG(128)
#undef CLIPB200_CFG
  return cudaSuccess;
}

inline int gemm_pick_bn(int N) {
  // minimise padded columns; prefer the wider tile on ties (fewer A re-reads, lower smem bandwidth per MMA)
  const int cands[3] = {256, 192, 128};
  int best = 256;
  long long best_cost = -1;
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const long long padded = static_cast<long long>((N + bn - 1) / bn) * bn;
    if (best_cost < 0 || padded < best_cost) { best_cost = padded; best = bn; }
  }
  return best;
}

// A: [M,K] bf16 (lda elements), W: [N,K] bf16 (ldw elements).  force_bn: 0 = auto.
inline cudaError_t gemm_bf16(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw, int M,
                             int N, int K, int epi_mode, const GemmEpilogue& ep, int num_sms, cudaStream_t stream,
                             int force_bn = 0) {
  if (M <= 0 || N <= 0 || K <= 0) return cudaErrorInvalidValue;
  if ((K & 7) || (N & 7) || (lda & 7) || (ldw & 7) || (ep.ldc & 3)) return cudaErrorInvalidValue;
  const int bn = force_bn ? force_bn : gemm_pick_bn(N);
  CUtensorMap ta, tw;
  if (!make_tmap_bf16_2d(&ta, A, M, K, lda, GEMM_BM)) return cudaErrorUnknown;
  if (!make_tmap_bf16_2d(&tw, W, N, K, ldw, bn)) return cudaErrorUnknown;
#define CLIPB200_GEMM_CASE(BN_)                                                                     \
  if (bn == BN_) {                                                                                  \
    if (epi_mode == EPI_BF16) return gemm_launch_t<BN_, EPI_BF16>(ta, tw, M, N, K, ep, num_sms, stream);   \
    if (epi_mode == EPI_RESID) return gemm_launch_t<BN_, EPI_RESID>(ta, tw, M, N, K, ep, num_sms, stream); \
    return gemm_launch_t<BN_, EPI_F32>(ta, tw, M, N, K, ep, num_sms, stream);                       \
  }
  CLIPB200_GEMM_CASE(256)
  CLIPB200_GEMM_CASE(192)
  CLIPB200_GEMM_CASE(128)
#undef CLIPB200_GEMM_CASE
  return cudaErrorInvalidValue;
}

}  // namespace clipb200
