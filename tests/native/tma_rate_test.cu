// Microbenchmark for the next attention layout decision (profiles/r01g_attn_summary.md section 7): how fast does TMA
// deliver one 96-key V block of one head out of a [B, T, 3*H*hd] bf16 tensor (L2-resident re-reads, as in the attention
// kernel) when the block is cut into boxes of different widths?
//   A  1 x {64 cols, 96 rows} 128B-swizzle + 1 x {8 cols, 96 rows} no swizzle      (today's main tile + remainder plane)
//   B  5 x {16 cols, 96 rows} 32B-swizzle                                          (one MN-major N=80 operand)
//   D  9 x { 8 cols, 96 rows} no swizzle                                           (all chunk planes)
//   E  2 x {64 cols, 96 rows} 128B-swizzle                                         (second swizzle atom for columns 64..127)
// Two CTAs per SM, one issuing thread each, four tiles in flight per CTA.  Run on a B200 only; prints a table.
#include <stdio.h>
#include <stdlib.h>

#include "../../clip_embedder_rs_b200/csrc/attn_sm100.cuh"

using namespace clipb200;

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(2);                                                                                 \
    }                                                                                          \
  } while (0)

constexpr int ROWS = 96, SLOTS = 4, SLOT_BYTES = 2 * ROWS * 128;   // room for variant E
constexpr int SMEM = SLOTS * SLOT_BYTES + 1024 + 64;

static bool make_map(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t T, uint64_t B, uint32_t box_cols,
                     CUtensorMapSwizzle sw) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return false;
  cuuint64_t dims[3] = {cols, T, B};
  cuuint64_t strides[2] = {cols * 2, cols * 2 * T};
  cuuint32_t box[3] = {box_cols, ROWS, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// n_box boxes of box_cols columns (box_bytes each) per tile, plus n_box2 boxes from the second map
__global__ void __launch_bounds__(32, 2)
tma_rate_kernel(const __grid_constant__ CUtensorMap tm1, const __grid_constant__ CUtensorMap tm2, int n_box1,
                int cols1, int bytes1, int n_box2, int cols2, int bytes2, int H, int hd, int T, int B, int tiles_per_cta,
                long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SLOTS * SLOT_BYTES);
  if (threadIdx.x == 0) {
    for (int s = 0; s < SLOTS; ++s) ptx::mbar_init(&bar[s], 1);
    ptx::fence_mbar_init();
    const int blocks_per_seq = T / ROWS;
    // working set of 2048 blocks (~60 MB of touched lines): L2-resident after the first pass, like the attention
    // kernel's K/V re-reads
    const int total = B * H * blocks_per_seq < 2048 ? B * H * blocks_per_seq : 2048;
    const uint32_t tx = static_cast<uint32_t>(n_box1 * bytes1 + n_box2 * bytes2);
    const long long t0 = clock64();
    for (int i = 0; i < tiles_per_cta + SLOTS; ++i) {
      if (i >= SLOTS) ptx::mbar_wait(&bar[i % SLOTS], ((i / SLOTS) - 1) & 1);   // the tile issued SLOTS iterations ago
      if (i < tiles_per_cta) {
        // walk (b, h, block) like the attention kernel's items: neighbouring CTAs read neighbouring heads / blocks
        const int item = (blockIdx.x + i * gridDim.x) % total;
        const int j = item % blocks_per_seq, bh = item / blocks_per_seq, h = bh % H, b = bh / H;
        const int col_v = 2 * H * hd + h * hd;
        uint8_t* dst = smem + (i % SLOTS) * SLOT_BYTES;
        ptx::mbar_arrive_expect_tx(&bar[i % SLOTS], tx);
        for (int k = 0; k < n_box1; ++k)
          attn::tma_load_3d(&tm1, &bar[i % SLOTS], dst + k * bytes1, col_v + k * cols1, j * ROWS, b);
        for (int k = 0; k < n_box2; ++k)
          attn::tma_load_3d(&tm2, &bar[i % SLOTS], dst + n_box1 * bytes1 + k * bytes2, col_v + n_box1 * cols1 + k * cols2,
                            j * ROWS, b);
      }
    }
    if (blockIdx.x == 0) out[0] = clock64() - t0;
  }
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int B = 128, T = 576, H = 16, hd = 72;
  const size_t n = (size_t)B * T * 3 * H * hd;
  __nv_bfloat16* qkv;
  CK(cudaMalloc(&qkv, n * 2));
  CK(cudaMemset(qkv, 0, n * 2));
  long long* d_out;
  CK(cudaMalloc(&d_out, 8));
  CK(cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  const uint64_t cols = 3ull * H * hd;
  CUtensorMap m64, m16, m8;
  if (!make_map(&m64, qkv, cols, T, B, 64, CU_TENSOR_MAP_SWIZZLE_128B) ||
      !make_map(&m16, qkv, cols, T, B, 16, CU_TENSOR_MAP_SWIZZLE_32B) ||
      !make_map(&m8, qkv, cols, T, B, 8, CU_TENSOR_MAP_SWIZZLE_NONE)) {
    printf("tensor map encode failed\n");
    return 2;
  }
  struct V { const char* name; const CUtensorMap* a; int na, ca, ba; const CUtensorMap* b; int nb, cb, bb; };
  const V vs[] = {
      {"A 64sw128 + 8 plain", &m64, 1, 64, ROWS * 128, &m8, 1, 8, ROWS * 16},
      {"B 5 x 16 sw32", &m16, 5, 16, ROWS * 32, &m8, 0, 8, ROWS * 16},
      {"D 9 x 8 plain", &m8, 9, 8, ROWS * 16, &m8, 0, 8, ROWS * 16},
      {"E 2 x 64 sw128", &m64, 2, 64, ROWS * 128, &m8, 0, 8, ROWS * 16},
  };
  const int grid = 2 * prop.multiProcessorCount, tiles = 400;
  printf("TMA delivery of one 96-key V block (hd 72) per tile, %d CTAs, %d tiles each, %d in flight per CTA\n", grid, tiles,
         SLOTS);
  printf("%-22s %-12s %-16s %-14s %-12s\n", "variant", "ms", "cycles/tile/CTA", "useful GB/s", "moved GB/s");
  for (const V& v : vs) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      tma_rate_kernel<<<grid, 32, SMEM>>>(*v.a, *v.b, v.na, v.ca, v.ba, v.nb, v.cb, v.bb, H, hd, T, B, tiles, d_out);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
    }
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long cyc;
    CK(cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost));
    const double useful = (double)grid * tiles * ROWS * hd * 2, moved = (double)grid * tiles * (v.na * v.ba + v.nb * v.bb);
    printf("%-22s %-12.3f %-16.0f %-14.0f %-12.0f\n", v.name, ms, (double)cyc / tiles, useful / ms * 1e-6, moved / ms * 1e-6);
  }
  printf("TMA RATE TEST DONE\n");
  return 0;
}
