// Standalone check of the tcgen05 flash-attention kernel against the mma.sync kernel (itself validated end to end
// against the CPU oracle) and against a plain fp32 SIMT reference on a few rows.  Run on a B200 only.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../clip_embedder_rs_b200/csrc/attn_sm100.cuh"
#include "../../clip_embedder_rs_b200/csrc/attn_short_sm100.cuh"
#include "../../clip_embedder_rs_b200/csrc/kernels.cuh"

using namespace clipb200;

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(2);                                                                                 \
    }                                                                                          \
  } while (0)

__global__ void fill_bf16(__nv_bfloat16* p, size_t n, uint32_t seed, float scale) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t z = (i + 1) * 0x9E3779B97F4A7C15ull + seed;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  p[i] = __float2bfloat16(((float)(z & 0xFFFFFF) / 16777216.0f - 0.5f) * scale);
}

// fp32 reference for one (b, h, q) row per thread
__global__ void ref_attn(const __nv_bfloat16* qkv, float* out, int B, int T, int H, int hd, int causal,
                         const int* qsel, int nsel) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H * nsel) return;
  int qi = qsel[idx % nsel];
  int h = (idx / nsel) % H, b = idx / (nsel * H);
  const long long D3 = 3ll * H * hd;
  const __nv_bfloat16* q = qkv + ((long long)b * T + qi) * D3 + h * hd;
  float scale = rsqrtf((float)hd);
  float m = -INFINITY, l = 0.f;
  float acc[128];
  for (int d = 0; d < hd; ++d) acc[d] = 0.f;
  int kmax = causal ? qi + 1 : T;
  for (int k = 0; k < kmax; ++k) {
    const __nv_bfloat16* kp = qkv + ((long long)b * T + k) * D3 + H * hd + h * hd;
    float s = 0.f;
    for (int d = 0; d < hd; ++d) s += __bfloat162float(q[d]) * __bfloat162float(kp[d]);
    s *= scale;
    float mn = fmaxf(m, s);
    float a = expf(m - mn), pv = expf(s - mn);
    const __nv_bfloat16* vp = kp + H * hd;
    for (int d = 0; d < hd; ++d) acc[d] = acc[d] * a + pv * __bfloat162float(vp[d]);
    l = l * a + pv;
    m = mn;
  }
  for (int d = 0; d < hd; ++d) out[(long long)idx * hd + d] = acc[d] / l;
}

// vt[b][h*hd + d][t] = v[b][t][h*hd + d] (what the qkv GEMM's EPI_QKVT epilogue writes)
__global__ void transpose_v(const __nv_bfloat16* qkv, __nv_bfloat16* vt, int B, int T, int H, int hd, int ld) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long D = (long long)H * hd;
  if (idx >= (long long)B * T * D) return;
  const int c = (int)(idx % D);
  const long long bt = idx / D;
  const int t = (int)(bt % T), b = (int)(bt / T);
  vt[((long long)b * D + c) * ld + t] = qkv[bt * 3 * D + 2 * D + c];
}

static int run_case(int B, int T, int H, int hd, bool causal, bool time_it, int num_sms) {
  const size_t nq = (size_t)B * T * 3 * H * hd, no = (size_t)B * T * H * hd;
  __nv_bfloat16 *qkv, *o1, *o2;
  CK(cudaMalloc(&qkv, nq * 2));
  CK(cudaMalloc(&o1, no * 2));
  CK(cudaMalloc(&o2, no * 2));
  fill_bf16<<<(unsigned)((nq + 255) / 256), 256>>>(qkv, nq, 7, 4.0f);
  CK(cudaMemset(o1, 0, no * 2));
  CK(cudaMemset(o2, 0x7f, no * 2));
  // transposed V for the single-PV-instruction kernel; poisoned first so that unwritten padding shows up
  const int ld = attn::attn_vt_ld(T);
  const size_t nvt = (size_t)B * H * hd * ld;
  __nv_bfloat16* vt;
  CK(cudaMalloc(&vt, nvt * 2));
  CK(cudaMemset(vt, 0x7f, nvt * 2));
  transpose_v<<<(unsigned)(((size_t)B * T * H * hd + 255) / 256), 256>>>(qkv, vt, B, T, H, hd, ld);
  static const bool use_vt = getenv("ATTN_NO_VT") == nullptr;
  auto tc = [&](__nv_bfloat16* o) {
    return use_vt ? attn_tcgen05_vt(qkv, vt, o, B, T, H, hd, causal, num_sms, 0) : attn_auto(qkv, o, B, T, H, hd, causal, num_sms, 0);
  };
  CK(launch_flash_attention(qkv, o1, B, T, H, hd, causal, 0));
  CK(tc(o2));
  CK(cudaDeviceSynchronize());
  std::vector<__nv_bfloat16> h1(no), h2(no);
  CK(cudaMemcpy(h1.data(), o1, no * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h2.data(), o2, no * 2, cudaMemcpyDeviceToHost));
  double max_d = 0;
  long long bad = 0, nan = 0;
  for (size_t i = 0; i < no; ++i) {
    float a = __bfloat162float(h1[i]), b = __bfloat162float(h2[i]);
    if (!(b == b)) { ++nan; continue; }
    double d = fabs((double)a - b);
    if (d > max_d) max_d = d;
    if (d > 0.02 + 0.02 * fabs(a)) {
      if (bad < 5) printf("   mismatch i=%zu (row %zu col %zu) mma.sync=%f tcgen05=%f\n", i, i / (H * hd), i % (H * hd), a, b);
      ++bad;
    }
  }
  // fp32 reference on a few query rows
  std::vector<int> qsel = {0, T > 1 ? 1 : 0, T / 2, T - 1};
  if (T > 130) { qsel.push_back(127); qsel.push_back(128); qsel.push_back(T - 64); }
  int* dq; float* dref;
  const int nsel = (int)qsel.size();
  CK(cudaMalloc(&dq, nsel * 4));
  CK(cudaMalloc(&dref, (size_t)B * H * nsel * hd * 4));
  CK(cudaMemcpy(dq, qsel.data(), nsel * 4, cudaMemcpyHostToDevice));
  ref_attn<<<(B * H * nsel + 63) / 64, 64>>>(qkv, dref, B, T, H, hd, causal ? 1 : 0, dq, nsel);
  CK(cudaDeviceSynchronize());
  std::vector<float> href((size_t)B * H * nsel * hd);
  CK(cudaMemcpy(href.data(), dref, href.size() * 4, cudaMemcpyDeviceToHost));
  double max_ref = 0;
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h)
      for (int s = 0; s < nsel; ++s)
        for (int d = 0; d < hd; ++d) {
          float want = href[(((size_t)b * H + h) * nsel + s) * hd + d];
          float got = __bfloat162float(h2[((size_t)b * T + qsel[s]) * H * hd + h * hd + d]);
          double dd = fabs((double)want - got);
          if (dd > max_ref) max_ref = dd;
          if (dd > 0.03 + 0.02 * fabs(want)) ++bad;
        }
  double ms1 = 0, ms2 = 0;
  if (time_it) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) {
      for (int i = 0; i < 3; ++i) {
        if (w == 0) CK(launch_flash_attention(qkv, o1, B, T, H, hd, causal, 0));
        else CK(tc(o2));
      }
      CK(cudaEventRecord(e0));
      for (int i = 0; i < 10; ++i) {
        if (w == 0) CK(launch_flash_attention(qkv, o1, B, T, H, hd, causal, 0));
        else CK(tc(o2));
      }
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float t; CK(cudaEventElapsedTime(&t, e0, e1));
      (w == 0 ? ms1 : ms2) = t / 10;
    }
  }
  const double flops = 4.0 * B * H * (double)T * T * hd * (causal ? 0.5 : 1.0);
  printf("%s B=%d T=%d H=%d hd=%d causal=%d max|tc-mma|=%.4g max|tc-fp32|=%.4g nan=%lld bad=%lld", (bad || nan) ? "FAIL" : "ok  ",
         B, T, H, hd, causal ? 1 : 0, max_d, max_ref, nan, bad);
  if (time_it) printf("  mma.sync %.3f ms (%.0f TF)  tcgen05%s %.3f ms (%.0f TF)", ms1, flops / ms1 * 1e-9, use_vt ? "+Vt" : "", ms2, flops / ms2 * 1e-9);
  printf("\n");
  fflush(stdout);
  cudaFree(qkv); cudaFree(o1); cudaFree(o2); cudaFree(dq); cudaFree(dref); cudaFree(vt);
  return (bad || nan) ? 1 : 0;
}

int main(int argc, char** argv) {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int num_sms = getenv("ATTN_HALF_GRID") ? prop.multiProcessorCount / 2 : prop.multiProcessorCount;
  CK(flash_attention_configure_device());
  CK(attn_configure_all());
  int fails = 0;
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  int idx = 0;
  auto run = [&](int B, int T, int H, int hd, bool causal, bool time_it) {
    if (only < 0 || only == idx) fails += run_case(B, T, H, hd, causal, time_it, num_sms);
    ++idx;
  };
  run(1, 96, 1, 64, false, false);     // 0: one item, one kv block, main part only
  run(1, 128, 1, 64, false, false);    // 1: kv tail masking (2 blocks, second partial)
  run(1, 96, 1, 72, false, false);     // 2: remainder planes (8 real + 8 zero)
  run(1, 96, 1, 80, false, false);     // 3
  run(1, 96, 1, 96, false, false);     // 4
  run(2, 576, 16, 72, false, false);   // 5: SO400M shape, multi item, persistent loop
  run(3, 50, 12, 64, false, false);    // 6: ViT-B/32
  run(2, 730, 16, 80, false, false);   // 7: ViT-H/14
  run(4, 77, 8, 64, true, false);      // 8: CLIP text (causal)
  run(2, 300, 4, 64, true, false);     // 9: causal, several q tiles
  run(2, 576, 16, 96, false, false);   // 10: giant-opt
  run(128, 576, 16, 72, false, true);  // 11: perf, SO400M micro-batch
  run(64, 576, 16, 96, false, true);   // 12
  run(128, 576, 18, 64, false, true);  // 13: no remainder planes (same total width as 16 x 72)
  run(128, 576, 14, 80, false, true);  // 14: two real remainder planes
  run(32, 2304, 16, 72, false, true);  // 15: 24 key blocks per item (per-block vs per-item cost, with 16)
  run(256, 288, 16, 72, false, true);  // 16: 3 key blocks per item
  run(2048, 77, 16, 64, true, true);   // 17: DFN5B text micro-batch: one causal key block per item
  run(2048, 77, 8, 64, true, true);    // 18: ViT-B/32 text width
  run(3, 1, 2, 64, true, false);       // 19..24: short-sequence kernel edges (one token; one / two / three active warps; T = 80)
  run(2, 33, 4, 64, true, false);
  run(2, 64, 2, 64, false, false);
  run(5, 65, 3, 64, false, false);
  run(2, 80, 2, 64, true, false);
  run(300, 80, 1, 64, false, false);   // more items than resident CTAs: the persistent loop and its parities
#ifdef CLIPB200_ATTN_TIMING
  {
    unsigned long long h[16];
    CK(cudaMemcpyFromSymbol(h, attn::g_attn_wait, sizeof(h)));
    const char* names[13] = {"prod q_empty", "prod k_empty", "prod v_empty", "mma k_full", "mma s_empty", "mma q_full",
                             "mma v_full", "mma p_full", "smx s_full", "smx pv_done", "smx pv_done(epi)",
                             "mma QK issue", "mma PV issue"};
    for (int i = 0; i < 13; ++i) printf("  wait[%-16s] = %llu cycles\n", names[i], h[i]);
  }
#endif
  printf("%s (%d failing cases)\n", fails ? "ATTN TEST FAILED" : "ATTN TEST PASSED", fails);
  return fails ? 1 : 0;
}
