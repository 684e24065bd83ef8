// How fast can caller-owned (pageable) photo buffers reach the GPU?  The photo path (Engine::VisionEmbedRgb8Var) stages
// them through pinned memory with up to 16 copy threads; the alternative is to pin the caller's pages in place
// (cudaHostRegister), DMA straight out of them and unpin.  This measures both on the box at hand.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
  const size_t bytes = 30u << 20;   // one photo
  const int n = 32;
  std::vector<unsigned char*> bufs(n);
  for (int i = 0; i < n; ++i) { bufs[i] = static_cast<unsigned char*>(malloc(bytes)); memset(bufs[i], i + 1, bytes); }
  unsigned char* dev; CK(cudaMalloc(&dev, bytes * 2));
  unsigned char* pinned; CK(cudaMallocHost(&pinned, bytes * 2));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  CK(cudaMemcpyAsync(dev, pinned, bytes, cudaMemcpyHostToDevice, st)); CK(cudaStreamSynchronize(st));
  // (a) pinned staging, one copy thread, then H2D
  double t0 = now();
  for (int i = 0; i < n; ++i) {
    memcpy(pinned + (i & 1) * bytes, bufs[i], bytes);
    CK(cudaMemcpyAsync(dev + (i & 1) * bytes, pinned + (i & 1) * bytes, bytes, cudaMemcpyHostToDevice, st));
    if (i & 1) CK(cudaStreamSynchronize(st));
  }
  CK(cudaStreamSynchronize(st));
  double t1 = now();
  printf("staged through pinned memory, 1 copy thread : %.1f GB/s\n", n * bytes / (t1 - t0) * 1e-9);
  // (b) register in place, DMA, unregister — sequential
  double treg = 0, tcopy = 0, tunreg = 0;
  for (int i = 0; i < n; ++i) {
    double a = now();
    CK(cudaHostRegister(bufs[i], bytes, cudaHostRegisterDefault));
    double b = now();
    CK(cudaMemcpyAsync(dev, bufs[i], bytes, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    double c = now();
    CK(cudaHostUnregister(bufs[i]));
    double d = now();
    treg += b - a; tcopy += c - b; tunreg += d - c;
  }
  printf("register in place: register %.2f ms (%.1f GB/s), DMA %.2f ms (%.1f GB/s), unregister %.2f ms per 30 MB photo\n",
         treg / n * 1e3, bytes / (treg / n) * 1e-9, tcopy / n * 1e3, bytes / (tcopy / n) * 1e-9, tunreg / n * 1e3);
  // (c) register / unregister from 1, 4 and 8 threads at once (does pinning scale?)
  for (int th : {1, 4, 8}) {
    double a = now();
    std::vector<std::thread> ts;
    for (int t = 0; t < th; ++t)
      ts.emplace_back([&, t] {
        for (int i = t; i < n; i += th) {
          if (cudaHostRegister(bufs[i], bytes, cudaHostRegisterDefault) != cudaSuccess) { printf("register failed\n"); exit(2); }
        }
      });
    for (auto& t : ts) t.join();
    double b = now();
    for (int i = 0; i < n; ++i) CK(cudaHostUnregister(bufs[i]));
    printf("cudaHostRegister from %d threads: %.1f GB/s aggregate\n", th, n * bytes / (b - a) * 1e-9);
  }
  // (d) plain pageable cudaMemcpyAsync
  t0 = now();
  for (int i = 0; i < n; ++i) CK(cudaMemcpyAsync(dev, bufs[i], bytes, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  t1 = now();
  printf("pageable cudaMemcpyAsync: %.1f GB/s\n", n * bytes / (t1 - t0) * 1e-9);
  printf("HOSTREG TEST DONE\n");
  return 0;
}
