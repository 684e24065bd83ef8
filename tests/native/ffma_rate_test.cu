// Issue rate of FFMA by operand form on sm_100a: three vector registers, or two vector registers + a uniform register
// (what nvcc emits when one multiplicand is a warp-uniform __constant__ value).  The depthwise 7x7 convolution is bound
// by this rate: with lanes across channels its weights are per-lane registers (first form); with lanes
// across pixels they are warp-uniform (second form).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__constant__ float W[64];
constexpr int ACC = 16, ITERS = 4096;

template <int FORM>
__global__ void __launch_bounds__(1024, 1) ffma_kernel(float* out, const float* in, long long* cycles) {
  float acc[ACC], x[8], wr[8];
#pragma unroll
  for (int j = 0; j < ACC; ++j) acc[j] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { x[j] = in[threadIdx.x * 8 + j]; wr[j] = in[1024 * 8 + threadIdx.x * 8 + j]; }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int j = 0; j < ACC; ++j)
        acc[j] = FORM == 0 ? fmaf(x[(j + k) & 7], wr[k], acc[j])      // three vector registers
                           : fmaf(x[(j + k) & 7], W[k], acc[j]);      // vector, uniform (constant bank), vector
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < ACC; ++j) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// The access pattern of the pixel-lane depthwise kernel: every warp of a CTA streams ITS OWN 196 taps (49 x 4 channels,
// tap-major table with a 4 * CH byte stride between taps) from the kernel-parameter constant bank, 8 FFMA per tap.
// LAYOUT 0: [49][CH] tap-major (stride between a warp's taps);  LAYOUT 1: [CH / 4][49][4] (a warp's taps contiguous).
template <int CH>
struct Taps { float w[49 * CH]; };
template <int CH, int LAYOUT, int WG>
__device__ __forceinline__ void taps_body(const Taps<CH>& prm, int lc0, const float (&x)[8], float (&acc)[8][4]) {
#pragma unroll 1
  for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) {
      const int t = ky * 7 + kx;
      float w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        w[j] = LAYOUT == 0 ? prm.w[t * CH + lc0 + 4 * WG + j] : prm.w[((lc0 / 4 + WG) * 49 + t) * 4 + j];
#pragma unroll
      for (int p = 0; p < 8; ++p)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[p][j] = fmaf(x[(p + kx) & 7], w[j], acc[p][j]);
    }
  }
}
template <int CH, int LAYOUT>
__global__ void __launch_bounds__(256, 2) taps_kernel(const __grid_constant__ Taps<CH> prm, float* out, const float* in, long long* cycles, int tiles) {
  float acc[8][4], x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = in[threadIdx.x * 8 + j];
  const int warp = threadIdx.x >> 5;
  __syncthreads();
  const long long t0 = clock64();
  for (int tile = 0; tile < tiles; ++tile) {
    const int lc0 = ((blockIdx.x + tile) % (CH / 32)) * 32;   // a different channel block every tile, like the real grid
#pragma unroll
    for (int p = 0; p < 8; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[p][j] = 0.f;
    switch (warp) {
      case 0: taps_body<CH, LAYOUT, 0>(prm, lc0, x, acc); break;
      case 1: taps_body<CH, LAYOUT, 1>(prm, lc0, x, acc); break;
      case 2: taps_body<CH, LAYOUT, 2>(prm, lc0, x, acc); break;
      case 3: taps_body<CH, LAYOUT, 3>(prm, lc0, x, acc); break;
      case 4: taps_body<CH, LAYOUT, 4>(prm, lc0, x, acc); break;
      case 5: taps_body<CH, LAYOUT, 5>(prm, lc0, x, acc); break;
      case 6: taps_body<CH, LAYOUT, 6>(prm, lc0, x, acc); break;
      default: taps_body<CH, LAYOUT, 7>(prm, lc0, x, acc); break;
    }
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < 8; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) s += acc[p][j];
    x[tile & 7] += s * 1e-30f;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x[0] + x[3];
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int CH, int LAYOUT>
void run_taps(float* out, const float* in, long long* cyc) {
  static Taps<CH> h;
  for (int i = 0; i < 49 * CH; ++i) h.w[i] = 1e-3f * (i % 97);
  const int tiles = 64;
  for (int rep = 0; rep < 2; ++rep) {
    taps_kernel<CH, LAYOUT><<<296, 256>>>(h, out, in, cyc, tiles);
    CK(cudaDeviceSynchronize());
  }
  long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
  // per scheduler: 2 CTAs x 2 warps, each tiles * 1568 FFMA
  const double per_sched = 4.0 * tiles * 1568;
  printf("taps from the parameter bank, %d channels (%d KB), layout %s: %.3f FFMA warp-instructions per clock per scheduler (%.0f FMA/clk/SM)\n",
         CH, 49 * CH * 4 / 1024, LAYOUT == 0 ? "[49][CH]" : "[CH/4][49][4]", per_sched / c, per_sched / c * 128);
}

int main() {
  float *in, *out; long long* cyc;
  CK(cudaMalloc(&in, 2 * 1024 * 8 * 4)); CK(cudaMemset(in, 0, 2 * 1024 * 8 * 4));
  CK(cudaMalloc(&out, 148 * 1024 * 4)); CK(cudaMalloc(&cyc, 8));
  float hw[64]; for (int i = 0; i < 64; ++i) hw[i] = 0.001f * i;
  CK(cudaMemcpyToSymbol(W, hw, sizeof(hw)));
  for (int form = 0; form < 2; ++form) {
    for (int rep = 0; rep < 2; ++rep) {
      if (form == 0) ffma_kernel<0><<<148, 1024>>>(out, in, cyc); else ffma_kernel<1><<<148, 1024>>>(out, in, cyc);
      CK(cudaDeviceSynchronize());
    }
    long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    // 1024 threads = 32 warps = 8 per scheduler; per warp ITERS * 8 * ACC FFMA instructions
    const double per_sched = 8.0 * ITERS * 8 * ACC;
    printf("%s: %lld cycles, %.3f FFMA warp-instructions per clock per scheduler (%.0f FMA/clk/SM)\n",
           form == 0 ? "FFMA R, R, R, R " : "FFMA R, R, UR, R", h, per_sched / h, per_sched / h * 4 * 32);
  }
  run_taps<32, 0>(out, in, cyc);
  run_taps<160, 0>(out, in, cyc);
  run_taps<160, 1>(out, in, cyc);
  printf("FFMA RATE TEST DONE\n");
  return 0;
}
