// Microbenchmark behind the attention kernel's design questions (profiles/r01g_attn_summary.md): what does one
// tcgen05.mma cost when it is chained on the same accumulator (the k-steps of S = Q K^T and O += P V) compared with
// independent accumulators, how long is issue -> tcgen05.commit -> mbarrier, and what does one mbarrier hand-off between
// two warps cost.  One CTA, data is zeros (timing does not depend on it).  Run on a B200 only; prints a table.
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

constexpr int A_BYTES = 128 * 128;   // 128 rows x 64 bf16, 128-byte-swizzled K-major tile
constexpr int B_BYTES = 256 * 128;   // up to N = 256 rows
constexpr int SMEM = A_BYTES + B_BYTES + 1024 + 256;

// mode 0: SS (A and B from shared memory); mode 1: TS (A from TMEM columns 448.., like P in the attention kernel)
// `chains` independent accumulators are used round-robin; every MMA accumulates (K = 16 per instruction).
__global__ void __launch_bounds__(64, 1) mma_chain_kernel(int n, int n_mma, int chains, int mode, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_BYTES);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 4);
  // lane-0 broadcasts keep the issue loop's operands in uniform registers (no R2UR per MMA), as in the kernels
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  ptx::fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar[0], 1);
    ptx::mbar_init(&bar[1], 1);
    ptx::mbar_init(&bar[2], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc<512>(tmem_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  if (warp == 0) {
    const uint32_t idesc = attn::make_idesc(128, n, 0);
    const uint64_t da = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem));
    const uint64_t db = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem + A_BYTES));
    const int stride = n < 32 ? 32 : n;   // accumulator c lives at columns c * stride (4 x 96 fits below 448)
    // warm-up pass, then the timed pass
    for (int pass = 0; pass < 2; ++pass) {
      __syncwarp();
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const uint32_t d = tmem + static_cast<uint32_t>((i & (chains - 1)) * stride);   // chains is a power of two
        const uint64_t k = static_cast<uint64_t>(2 * (i & 3));   // the four K = 16 slices of the 64-wide tile
        if (mode == 0) ptx::umma_bf16_ss_w(d, da + k, db + k, idesc, 1u);
        else ptx::umma_bf16_ts_w(d, tmem + 448u + static_cast<uint32_t>(8 * (i & 3)), db + k, idesc, 1u);
      }
      const long long t1 = clock64();
      ptx::umma_commit_w(&bar[0]);
      ptx::mbar_wait(&bar[0], pass & 1);
      const long long t2 = clock64();
      if (pass == 1 && lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc<512>(tmem); }
}

// Two warps bounce one phase back and forth through two mbarriers `rounds` times: cycles per one-way hand-off.
// wait_mode 0: the library's mbar_wait (try_wait, suspended with a time hint), 1: test_wait spin
__global__ void __launch_bounds__(64, 1) hop_kernel(int rounds, int wait_mode, int lanes_arrive, long long* out) {
  __shared__ uint64_t bar[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar[0], lanes_arrive);
    ptx::mbar_init(&bar[1], lanes_arrive);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  auto wait = [&](uint64_t* b, uint32_t par) {
    if (wait_mode == 0) ptx::mbar_wait(b, par);
    else while (!ptx::mbar_test_wait(b, par)) {}
  };
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    if (warp == 0) {
      if (lane < lanes_arrive) ptx::mbar_arrive(&bar[0]);
      wait(&bar[1], r & 1);
    } else {
      wait(&bar[0], r & 1);
      if (lane < lanes_arrive) ptx::mbar_arrive(&bar[1]);
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = (t1 - t0) / (2 * rounds);
}

int main() {
  CK(cudaSetDevice(0));
  long long* d_out;
  CK(cudaMalloc(&d_out, 64));
  CK(cudaFuncSetAttribute(mma_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  long long h[2];
  printf("tcgen05.mma M=128 K=16 bf16, one CTA, 64 MMAs: cycles per MMA (issue loop | until commit observed)\n");
  printf("%-6s %-5s %-8s %-18s %-18s\n", "mode", "N", "chains", "issue/MMA", "complete/MMA");
  const int ns[] = {16, 32, 64, 80, 96, 128, 256};
  for (int mode = 0; mode < 2; ++mode)
    for (int n : ns)
      for (int chains = 1; chains <= 4; chains *= 2) {
        if (chains * (n < 32 ? 32 : n) > 448) continue;
        mma_chain_kernel<<<1, 64, SMEM>>>(n, 64, chains, mode, d_out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
        printf("%-6s %-5d %-8d %-18.1f %-18.1f\n", mode == 0 ? "SS" : "TS", n, chains, h[0] / 64.0, h[1] / 64.0);
      }
  printf("\nsingle MMA (N=64): issue -> commit -> mbarrier observed by the issuing warp\n");
  for (int n_mma : {1, 2, 4, 8}) {
    mma_chain_kernel<<<1, 64, SMEM>>>(64, n_mma, 1, 0, d_out);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
    printf("  %d dependent MMAs: issue %lld cycles, complete %lld cycles\n", n_mma, h[0], h[1]);
  }
  printf("\nmbarrier hand-off between two warps (cycles per one-way hop)\n");
  for (int wm = 0; wm < 2; ++wm)
    for (int lanes : {1, 32}) {
      hop_kernel<<<1, 64>>>(2000, wm, lanes, d_out);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, d_out, 8, cudaMemcpyDeviceToHost));
      printf("  %-28s arrivals per phase %-3d: %lld cycles\n", wm == 0 ? "try_wait (suspended)" : "test_wait spin", lanes, h[0]);
    }
  printf("MMA LATENCY TEST DONE\n");
  return 0;
}
