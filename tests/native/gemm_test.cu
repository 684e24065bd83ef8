// Standalone check of the tcgen05 GEMM against a plain SIMT reference (fp32 accumulate from the same
// bf16 inputs).  Prints max abs / rel error per shape and a timing for the large shapes.
// Build: see clip_embedder_rs_b200/csrc/Makefile (target gemm_test).  Run on a B200 only.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../clip_embedder_rs_b200/csrc/gemm_sm100.cuh"
#include "../../clip_embedder_rs_b200/csrc/fused_mlp_sm100.cuh"

using namespace clipb200;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

__global__ void fill_bf16(__nv_bfloat16* p, size_t n, uint32_t seed, float scale) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t z = (i + 1) * 0x9E3779B97F4A7C15ull + seed;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  float u = (float)(z & 0xFFFFFF) / 16777216.0f - 0.5f;
  p[i] = __float2bfloat16(u * scale);
}
__global__ void fill_f32(float* p, size_t n, uint32_t seed, float scale) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t z = (i + 1) * 0x9E3779B97F4A7C15ull + seed;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  p[i] = ((float)(z & 0xFFFFFF) / 16777216.0f - 0.5f) * scale;
}

// reference for selected rows: ref[i, n] = sum_k A[rows[i], k] * W[n, k]
__global__ void ref_rows(const __nv_bfloat16* A, const __nv_bfloat16* W, const int* rows, int nrows, int N, int K,
                         float* ref) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (n >= N || i >= nrows) return;
  const __nv_bfloat16* a = A + (size_t)rows[i] * K;
  const __nv_bfloat16* w = W + (size_t)n * K;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += __bfloat162float(a[k]) * __bfloat162float(w[k]);
  ref[(size_t)i * N + n] = acc;
}

static float host_act(float x, int act) {
  switch (act) {
    case ACT_QUICKGELU: return x / (1.0f + expf(-1.702f * x));
    case ACT_GELU_TANH: return 0.5f * x * (1.0f + tanhf(0.7978845608028654f * (x + 0.044715f * x * x * x)));
    case ACT_GELU_ERF: return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f));
    default: return x;
  }
}
static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

static int run_case(int M, int N, int K, int epi, int act, int force_bn, bool time_it, int num_sms, int ncta = 1) {
  __nv_bfloat16 *A, *W, *Cb;
  float *bias, *Cf, *Cf0, *gamma, *pos;
  const int rows_in = (epi == EPI_F32) ? 7 : 0, rows_out = 8, row_off = 1;  // CLS-style remap test
  const long long out_rows = (epi == EPI_F32) ? (long long)((M + rows_in - 1) / rows_in) * rows_out : M;
  CK(cudaMalloc(&A, (size_t)M * K * 2));
  CK(cudaMalloc(&W, (size_t)N * K * 2));
  CK(cudaMalloc(&Cb, (size_t)M * N * 2));
  CK(cudaMalloc(&Cf, (size_t)out_rows * N * 4));
  CK(cudaMalloc(&Cf0, (size_t)out_rows * N * 4));
  CK(cudaMalloc(&bias, (size_t)N * 4));
  CK(cudaMalloc(&gamma, (size_t)N * 4));
  CK(cudaMalloc(&pos, (size_t)rows_out * N * 4));
  auto blocks = [](size_t n) { return (unsigned)((n + 255) / 256); };
  fill_bf16<<<blocks((size_t)M * K), 256>>>(A, (size_t)M * K, 1, 2.0f);
  fill_bf16<<<blocks((size_t)N * K), 256>>>(W, (size_t)N * K, 2, 0.25f);
  fill_f32<<<blocks(N), 256>>>(bias, N, 3, 1.0f);
  fill_f32<<<blocks(N), 256>>>(gamma, N, 4, 2.0f);
  fill_f32<<<blocks((size_t)rows_out * N), 256>>>(pos, (size_t)rows_out * N, 5, 1.0f);
  fill_f32<<<blocks((size_t)out_rows * N), 256>>>(Cf0, (size_t)out_rows * N, 6, 1.0f);
  CK(cudaMemcpy(Cf, Cf0, (size_t)out_rows * N * 4, cudaMemcpyDeviceToDevice));
  CK(cudaMemset(Cb, 0, (size_t)M * N * 2));
  CK(cudaDeviceSynchronize());

  GemmEpilogue ep;
  ep.bias = bias;
  ep.act = act;
  ep.ldc = N;
  ep.out_bf16 = Cb;
  ep.out_f32 = Cf;
  if (epi == EPI_RESID) ep.gamma = gamma;
  if (epi == EPI_F32) { ep.pos = pos; ep.rows_in = rows_in; ep.rows_out = rows_out; ep.row_off = row_off; }
  CK(gemm_bf16(A, K, W, K, M, N, K, epi, ep, num_sms, 0, force_bn, ncta));
  CK(cudaDeviceSynchronize());

  // rows to verify
  std::vector<int> rows;
  if (M <= 512) { for (int r = 0; r < M; ++r) rows.push_back(r); }
  else {
    for (int r = 0; r < 160; ++r) rows.push_back(r);
    for (int r = M - 160; r < M; ++r) rows.push_back(r);
    for (int i = 0; i < 192; ++i) rows.push_back((int)(((long long)i * 2654435761ll) % M));
  }
  int nrows = (int)rows.size();
  int* drows; float* dref;
  CK(cudaMalloc(&drows, nrows * 4));
  CK(cudaMalloc(&dref, (size_t)nrows * N * 4));
  CK(cudaMemcpy(drows, rows.data(), nrows * 4, cudaMemcpyHostToDevice));
  ref_rows<<<dim3((N + 127) / 128, nrows), 128>>>(A, W, drows, nrows, N, K, dref);
  CK(cudaDeviceSynchronize());
  std::vector<float> ref((size_t)nrows * N), hb(N), hg(N), hpos((size_t)rows_out * N);
  CK(cudaMemcpy(ref.data(), dref, ref.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hb.data(), bias, N * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hg.data(), gamma, N * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hpos.data(), pos, hpos.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<__nv_bfloat16> hcb;
  std::vector<float> hcf, hcf0;
  if (epi == EPI_BF16) { hcb.resize((size_t)M * N); CK(cudaMemcpy(hcb.data(), Cb, hcb.size() * 2, cudaMemcpyDeviceToHost)); }
  else {
    hcf.resize((size_t)out_rows * N); hcf0.resize((size_t)out_rows * N);
    CK(cudaMemcpy(hcf.data(), Cf, hcf.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hcf0.data(), Cf0, hcf0.size() * 4, cudaMemcpyDeviceToHost));
  }
  double max_abs = 0, max_rel = 0;
  long long bad = 0;
  for (int i = 0; i < nrows; ++i) {
    const int r = rows[i];
    for (int n = 0; n < N; ++n) {
      float want = ref[(size_t)i * N + n] + hb[n];
      float got;
      float tol;
      if (epi == EPI_BF16) {
        want = host_act(want, act);
        got = __bfloat162float(hcb[(size_t)r * N + n]);
        tol = 0.02f + fabsf(want) * (1.0f / 128.0f);
      } else if (epi == EPI_RESID) {
        want = hcf0[(size_t)r * N + n] + hg[n] * want;
        got = hcf[(size_t)r * N + n];
        tol = 0.02f + fabsf(want) * 1e-3f;
      } else {
        const int b = r / rows_in, t = r % rows_in;
        const long long orow = (long long)b * rows_out + t + row_off;
        want = want + hpos[(size_t)(t + row_off) * N + n];
        got = hcf[(size_t)orow * N + n];
        tol = 0.02f + fabsf(want) * 1e-3f;
      }
      const double d = fabs((double)got - (double)want);
      if (d > max_abs) max_abs = d;
      const double rel = d / (fabs((double)want) + 1e-3);
      if (rel > max_rel) max_rel = rel;
      if (!(d <= tol)) {
        if (bad < 5) printf("   mismatch r=%d n=%d got=%f want=%f\n", r, n, got, want);
        ++bad;
      }
    }
  }
  // EPI_F32 remap: untouched rows (row 0 of every group of rows_out) must keep their original contents
  if (epi == EPI_F32) {
    for (long long g = 0; g < out_rows / rows_out; ++g)
      for (int n = 0; n < N; ++n)
        if (hcf[(size_t)(g * rows_out) * N + n] != hcf0[(size_t)(g * rows_out) * N + n]) { ++bad; }
  }
  double ms = 0;
  if (time_it) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) CK(gemm_bf16(A, K, W, K, M, N, K, epi, ep, num_sms, 0, force_bn, ncta));
    CK(cudaEventRecord(e0));
    const int iters = 10;
    for (int i = 0; i < iters; ++i) CK(gemm_bf16(A, K, W, K, M, N, K, epi, ep, num_sms, 0, force_bn, ncta));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float t; CK(cudaEventElapsedTime(&t, e0, e1));
    ms = t / iters;
  }
  const char* en = epi == EPI_BF16 ? "bf16" : (epi == EPI_RESID ? "resid" : "f32");
  printf("%s cta%d M=%d N=%d K=%d epi=%s act=%d bn=%d max_abs=%.4g max_rel=%.4g bad=%lld", bad ? "FAIL" : "ok  ", ncta, M, N, K,
         en, act, force_bn ? force_bn : gemm_pick_bn(N, K), max_abs, max_rel, bad);
  if (time_it) printf("  %.3f ms  %.1f TFLOP/s", ms, 2.0 * M * N * K / ms * 1e-9);
  printf("\n");
  fflush(stdout);
  cudaFree(A); cudaFree(W); cudaFree(Cb); cudaFree(Cf); cudaFree(Cf0); cudaFree(bias); cudaFree(gamma); cudaFree(pos);
  cudaFree(drows); cudaFree(dref);
  return bad ? 1 : 0;
}

// EPI_QKVT: the q | k columns must equal the plain bf16 epilogue's, the v columns must appear transposed in
// vt[b][c][t] (ld = T rounded up to 8, padding untouched), bit for bit; optionally timed against the plain epilogue.
static int run_qkvt_case(int B, int T, int D, int K, bool time_it, int num_sms, int ncta) {
  const int M = B * T, N = 3 * D, ld = (T + 7) & ~7;
  __nv_bfloat16 *A, *W, *C0, *C1, *Vt;
  float* bias;
  CK(cudaMalloc(&A, (size_t)M * K * 2));
  CK(cudaMalloc(&W, (size_t)N * K * 2));
  CK(cudaMalloc(&C0, (size_t)M * N * 2));
  CK(cudaMalloc(&C1, (size_t)M * N * 2));
  CK(cudaMalloc(&Vt, (size_t)B * D * ld * 2));
  CK(cudaMalloc(&bias, (size_t)N * 4));
  auto blocks = [](size_t n) { return (unsigned)((n + 255) / 256); };
  fill_bf16<<<blocks((size_t)M * K), 256>>>(A, (size_t)M * K, 11, 2.0f);
  fill_bf16<<<blocks((size_t)N * K), 256>>>(W, (size_t)N * K, 12, 0.25f);
  fill_f32<<<blocks(N), 256>>>(bias, N, 13, 1.0f);
  CK(cudaMemset(C0, 0, (size_t)M * N * 2));
  CK(cudaMemset(C1, 0x11, (size_t)M * N * 2));
  CK(cudaMemset(Vt, 0x22, (size_t)B * D * ld * 2));
  GemmEpilogue e0, e1;
  e0.bias = bias; e0.ldc = N; e0.out_bf16 = C0;
  e1 = e0; e1.out_bf16 = C1; e1.out_vt = Vt; e1.vt_col0 = 2 * D; e1.vt_T = T; e1.vt_ld = ld; e1.vt_B = B;
  CK(gemm_bf16(A, K, W, K, M, N, K, EPI_BF16, e0, num_sms, 0, 0, ncta));
  CK(gemm_bf16(A, K, W, K, M, N, K, EPI_QKVT, e1, num_sms, 0, 0, ncta));
  CK(cudaDeviceSynchronize());
  std::vector<uint16_t> h0((size_t)M * N), h1((size_t)M * N), hv((size_t)B * D * ld);
  CK(cudaMemcpy(h0.data(), C0, h0.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h1.data(), C1, h1.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hv.data(), Vt, hv.size() * 2, cudaMemcpyDeviceToHost));
  long long bad = 0;
  for (int r = 0; r < M && bad < 10; ++r) {
    const int b = r / T, t = r % T;
    for (int c = 0; c < 2 * D; ++c)
      if (h0[(size_t)r * N + c] != h1[(size_t)r * N + c]) { if (bad < 5) printf("   q|k mismatch r=%d c=%d\n", r, c); ++bad; }
    for (int c = 0; c < D; ++c)
      if (hv[((size_t)b * D + c) * ld + t] != h0[(size_t)r * N + 2 * D + c]) {
        if (bad < 5) printf("   vt mismatch b=%d c=%d t=%d: %04x vs %04x\n", b, c, t, hv[((size_t)b * D + c) * ld + t], h0[(size_t)r * N + 2 * D + c]);
        ++bad;
      }
  }
  for (int b = 0; b < B && bad < 10; ++b)   // padding keys [T, ld) must stay untouched
    for (int c = 0; c < D; ++c)
      for (int t = T; t < ld; ++t)
        if (hv[((size_t)b * D + c) * ld + t] != 0x2222) ++bad;
  double ms[2] = {0, 0};
  if (time_it) {
    cudaEvent_t a, b2;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b2));
    for (int w = 0; w < 2; ++w) {
      for (int i = 0; i < 13; ++i) {
        if (i == 3) CK(cudaEventRecord(a));
        CK(gemm_bf16(A, K, W, K, M, N, K, w == 0 ? EPI_BF16 : EPI_QKVT, w == 0 ? e0 : e1, num_sms, 0, 0, ncta));
      }
      CK(cudaEventRecord(b2));
      CK(cudaEventSynchronize(b2));
      float t; CK(cudaEventElapsedTime(&t, a, b2));
      ms[w] = t / 10;
    }
  }
  printf("%s qkvt B=%d T=%d D=%d K=%d ncta=%d bad=%lld", bad ? "FAIL" : "ok  ", B, T, D, K, ncta, bad);
  if (time_it) printf("  bf16 %.3f ms (%.0f TF)  qkvt %.3f ms (%.0f TF)", ms[0], 2.0 * M * N * K / ms[0] * 1e-9, ms[1], 2.0 * M * N * K / ms[1] * 1e-9);
  printf("\n");
  fflush(stdout);
  cudaFree(A); cudaFree(W); cudaFree(C0); cudaFree(C1); cudaFree(Vt); cudaFree(bias);
  return bad ? 1 : 0;
}

// Fused ConvMlp tail (fused_mlp_sm100.cuh) against the two-launch path it replaces: fc1 + GELU as a bf16-store GEMM,
// fc2 + layer scale + residual as the fp32 reduce-add GEMM.  Both round the hidden activation to bf16 once, so the results
// differ only by fp32 summation order.
static int run_fused_mlp_case(int M, int C, int Hd, bool time_it, int num_sms) {
  __nv_bfloat16 *A, *W1, *W2, *Hbuf;
  float *b1, *b2, *gamma, *x0, *xa, *xb;
  CK(cudaMalloc(&A, (size_t)M * C * 2));
  CK(cudaMalloc(&W1, (size_t)Hd * C * 2));
  CK(cudaMalloc(&W2, (size_t)C * Hd * 2));
  CK(cudaMalloc(&Hbuf, (size_t)M * Hd * 2));
  CK(cudaMalloc(&b1, (size_t)Hd * 4));
  CK(cudaMalloc(&b2, (size_t)C * 4));
  CK(cudaMalloc(&gamma, (size_t)C * 4));
  CK(cudaMalloc(&x0, (size_t)M * C * 4));
  CK(cudaMalloc(&xa, (size_t)M * C * 4));
  CK(cudaMalloc(&xb, (size_t)M * C * 4));
  auto blocks = [](size_t n) { return (unsigned)((n + 255) / 256); };
  fill_bf16<<<blocks((size_t)M * C), 256>>>(A, (size_t)M * C, 21, 2.0f);
  fill_bf16<<<blocks((size_t)Hd * C), 256>>>(W1, (size_t)Hd * C, 22, 0.5f);
  fill_bf16<<<blocks((size_t)C * Hd), 256>>>(W2, (size_t)C * Hd, 23, 0.25f);
  fill_f32<<<blocks(Hd), 256>>>(b1, Hd, 24, 1.0f);
  fill_f32<<<blocks(C), 256>>>(b2, C, 25, 1.0f);
  fill_f32<<<blocks(C), 256>>>(gamma, C, 26, 2.0f);
  fill_f32<<<blocks((size_t)M * C), 256>>>(x0, (size_t)M * C, 27, 1.0f);
  CK(cudaMemcpy(xa, x0, (size_t)M * C * 4, cudaMemcpyDeviceToDevice));
  CK(cudaMemcpy(xb, x0, (size_t)M * C * 4, cudaMemcpyDeviceToDevice));
  auto unfused = [&](float* x) {
    GemmEpilogue e1;
    e1.bias = b1; e1.act = ACT_GELU_ERF; e1.ldc = Hd; e1.out_bf16 = Hbuf;
    CK(gemm_bf16(A, C, W1, C, M, Hd, C, EPI_BF16, e1, num_sms, 0));
    GemmEpilogue e2;
    e2.bias = b2; e2.gamma = gamma; e2.ldc = C; e2.out_f32 = x;
    CK(gemm_bf16(Hbuf, Hd, W2, Hd, M, C, Hd, EPI_RESID, e2, num_sms, 0));
  };
  auto fused = [&](float* x) { CK(fused_mlp(A, C, W1, C, b1, W2, Hd, b2, gamma, x, C, M, C, Hd, num_sms, 0)); };
  unfused(xa);
  fused(xb);
  CK(cudaDeviceSynchronize());
  std::vector<float> ha((size_t)M * C), hb((size_t)M * C), h0((size_t)M * C);
  CK(cudaMemcpy(ha.data(), xa, ha.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hb.data(), xb, hb.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h0.data(), x0, h0.size() * 4, cudaMemcpyDeviceToHost));
  double max_abs = 0, max_mag = 0;
  long long bad = 0, unchanged = 0;
  for (size_t i = 0; i < ha.size(); ++i) {
    const double d = fabs((double)ha[i] - hb[i]);
    max_abs = d > max_abs ? d : max_abs;
    max_mag = fabs(ha[i] - h0[i]) > max_mag ? fabs(ha[i] - h0[i]) : max_mag;
    if (!(hb[i] == hb[i]) || d > 2e-2 + 2e-3 * fabs(ha[i] - h0[i])) {
      if (bad < 5) printf("   mismatch row %zu col %zu: two launches %f fused %f (x0 %f)\n", i / C, i % C, ha[i], hb[i], h0[i]);
      ++bad;
    }
    if (hb[i] == h0[i]) ++unchanged;
  }
  if (unchanged > (long long)ha.size() / 100) { printf("   %lld outputs untouched by the fused kernel\n", unchanged); ++bad; }
  double ms[2] = {0, 0};
  if (time_it) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) {
      for (int i = 0; i < 13; ++i) {
        if (i == 3) CK(cudaEventRecord(e0));
        if (w == 0) unfused(xa); else fused(xb);
      }
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float t; CK(cudaEventElapsedTime(&t, e0, e1));
      ms[w] = t / 10;
    }
  }
#ifdef CLIPB200_FMLP_TIMING
  if (time_it) {
    unsigned long long* dbg;
    CK(cudaMalloc(&dbg, 32 * 8));
    CK(cudaMemset(dbg, 0, 32 * 8));
    clipb200::fmlp::timing_buffer() = dbg;
    fused(xb);
    CK(cudaDeviceSynchronize());
    clipb200::fmlp::timing_buffer() = nullptr;
    unsigned long long h[32];
    CK(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
    const double tiles = (double)h[9], chunks = (double)h[10];
    printf("   CTA 0, %.0f tiles, %.0f chunks; epilogue warp 0, clocks per chunk: wait S %.0f | tmem ld %.0f | GELU %.0f | wait H free %.0f | st + fence + arrive %.0f"
           "   per tile: wait O %.0f | final epilogue %.0f | tile-start barrier %.0f | total %.0f\n",
           tiles, chunks, h[0] / chunks, h[1] / chunks, h[2] / chunks, h[3] / chunks, h[4] / chunks, h[5] / tiles, h[6] / tiles, h[7] / tiles, h[8] / tiles);
    printf("   MMA warp, clocks per tile: wait A %.0f | wait S free %.0f | wait W1 %.0f | wait H %.0f | wait O free %.0f | wait W2 %.0f | total %.0f\n",
           h[16] / tiles, h[17] / tiles, h[18] / tiles, h[19] / tiles, h[20] / tiles, h[21] / tiles, h[24] / tiles);
    cudaFree(dbg);
  }
#endif
  const double flops = 4.0 * M * (double)C * Hd;
  printf("%s fused_mlp M=%d C=%d Hd=%d max|diff|=%.4g (update magnitude up to %.3g) bad=%lld", bad ? "FAIL" : "ok  ", M, C, Hd, max_abs, max_mag, bad);
  if (time_it) printf("  two launches %.1f us (%.0f TF)  fused %.1f us (%.0f TF)", ms[0] * 1e3, flops / ms[0] * 1e-9, ms[1] * 1e3, flops / ms[1] * 1e-9);
  printf("\n");
  fflush(stdout);
  cudaFree(A); cudaFree(W1); cudaFree(W2); cudaFree(Hbuf); cudaFree(b1); cudaFree(b2); cudaFree(gamma); cudaFree(x0); cudaFree(xa); cudaFree(xb);
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device %s sm_%d%d SMs=%d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  const int num_sms = prop.multiProcessorCount;
  CK(gemm_configure_device());
  int fails = 0;
  const bool quick = argc > 1 && atoi(argv[1]) == 1;
  if (argc > 1 && atoi(argv[1]) == 2) {  // CTA-pair tile-shape sweep
    for (int ncta = 1; ncta <= 2; ++ncta)
      for (int bn : {128, 192, 256}) fails += run_case(8192, 8064, 4096, EPI_BF16, ACT_NONE, bn, true, num_sms, ncta);
    return fails;
  }
  if (argc > 1 && atoi(argv[1]) == 3) {  // FastViT 1x1-conv shapes (MobileCLIP2-S2, 256 images): short K, epilogue-bound
    for (int ncta = 1; ncta <= 2; ++ncta) {
      for (int act : {ACT_NONE, ACT_GELU_ERF}) {
        fails += run_case(65536, 960, 320, EPI_BF16, act, 0, true, num_sms, ncta);      // stage 3 fc1
        fails += run_case(262144, 480, 160, EPI_BF16, act, 0, true, num_sms, ncta);     // stage 2 fc1
        fails += run_case(1048576, 240, 80, EPI_BF16, act, 0, true, num_sms, ncta);     // stage 1 fc1
      }
      fails += run_case(65536, 320, 960, EPI_RESID, ACT_NONE, 0, true, num_sms, ncta);  // stage 3 fc2
      fails += run_case(1048576, 80, 240, EPI_RESID, ACT_NONE, 0, true, num_sms, ncta); // stage 1 fc2
    }
    for (int bn : {128, 192, 256}) fails += run_case(65536, 960, 320, EPI_BF16, ACT_GELU_ERF, bn, true, num_sms, 2);
    return fails;
  }
  if (argc > 1 && atoi(argv[1]) == 4) {  // one short-K, epilogue-heavy launch for ncu (FastViT stage-1 fc1 + GELU)
    const int ncta = argc > 2 ? atoi(argv[2]) : 1;
    return run_case(1048576, 240, 80, EPI_BF16, ACT_GELU_ERF, 0, false, num_sms, ncta);
  }
  if (argc > 1 && atoi(argv[1]) == 8) {  // fused ConvMlp tail: small shapes with every tail, then the MobileCLIP2-S2 stage shapes timed
    CK(fused_mlp_configure_device());
    fails += run_fused_mlp_case(128, 80, 64, false, num_sms);        // one tile, one chunk
    fails += run_fused_mlp_case(128, 80, 240, false, num_sms);       // hidden tail (240 = 3 x 64 + 48)
    fails += run_fused_mlp_case(300, 160, 480, false, num_sms);      // row tail
    fails += run_fused_mlp_case(1000, 320, 960, false, num_sms);     // C > 256: two N halves, single-stage W ring
    fails += run_fused_mlp_case(777, 96, 384, false, num_sms);
    fails += run_fused_mlp_case(520, 256, 1024, false, num_sms);
    fails += run_fused_mlp_case(1048576, 80, 240, true, num_sms);    // S2 stage 1, 256 images
    fails += run_fused_mlp_case(262144, 160, 480, true, num_sms);    // stage 2
    fails += run_fused_mlp_case(65536, 320, 960, true, num_sms);     // stage 3
    printf("%s\n", fails ? "FUSED MLP TEST FAILED" : "FUSED MLP TEST PASSED");
    return fails;
  }
  if (argc > 1 && atoi(argv[1]) == 7) {  // transposing qkv epilogue: correctness on awkward shapes, then the SO400M qkv GEMM timed
    fails += run_qkvt_case(3, 32, 128, 128, false, num_sms, 1);     // one 32-token group per sequence, M tail (96 rows)
    fails += run_qkvt_case(5, 64, 128, 64, false, num_sms, 1);
    fails += run_qkvt_case(2, 576, 1152, 1152, false, num_sms, 1);
    fails += run_qkvt_case(7, 576, 1152, 1152, false, num_sms, 2);  // CTA pairs, M tail
    fails += run_qkvt_case(3, 736, 1280, 1280, false, num_sms, 2);
    fails += run_qkvt_case(256, 576, 1152, 1152, true, num_sms, 2);
    fails += run_qkvt_case(128, 576, 1536, 1536, true, num_sms, 2);
    printf("%s\n", fails ? "QKVT TEST FAILED" : "QKVT TEST PASSED");
    return fails;
  }
  if (argc > 1 && atoi(argv[1]) == 6) {  // what does the fp32 reduce-add epilogue cost?  same shapes, bf16 store instead
    const int M = 256 * 576;
    fails += run_case(M, 1152, 4304, EPI_RESID, ACT_NONE, 0, true, num_sms, 2);
    fails += run_case(M, 1152, 4304, EPI_BF16, ACT_NONE, 0, true, num_sms, 2);
    fails += run_case(M, 1152, 1152, EPI_RESID, ACT_NONE, 0, true, num_sms, 2);
    fails += run_case(M, 1152, 1152, EPI_BF16, ACT_NONE, 0, true, num_sms, 2);
    return fails;
  }
  if (argc > 1 && atoi(argv[1]) == 5) {  // tile-width sweep on the SO400M / DFN5B-text / giant-opt layer shapes (CTA pairs)
    const int M = 256 * 576;
    for (int bn : {128, 192, 256}) {
      fails += run_case(M, 3456, 1152, EPI_BF16, ACT_NONE, bn, true, num_sms, 2);
      fails += run_case(M, 1152, 1152, EPI_RESID, ACT_NONE, bn, true, num_sms, 2);
      fails += run_case(M, 4304, 1152, EPI_BF16, ACT_GELU_TANH, bn, true, num_sms, 2);
      fails += run_case(M, 1152, 4304, EPI_RESID, ACT_NONE, bn, true, num_sms, 2);
      fails += run_case(M, 3072, 1024, EPI_BF16, ACT_NONE, bn, true, num_sms, 2);
      fails += run_case(M, 1024, 4096, EPI_RESID, ACT_NONE, bn, true, num_sms, 2);
      fails += run_case(M, 4608, 1536, EPI_BF16, ACT_NONE, bn, true, num_sms, 2);
      fails += run_case(M, 1536, 6144, EPI_RESID, ACT_NONE, bn, true, num_sms, 2);
    }
    return fails;
  }
  // smallest cases first: one tile, one k-block
  fails += run_case(128, 256, 64, EPI_BF16, ACT_NONE, 256, false, num_sms);
  fails += run_case(128, 128, 64, EPI_BF16, ACT_NONE, 128, false, num_sms);
  fails += run_case(128, 192, 64, EPI_BF16, ACT_NONE, 192, false, num_sms);
  fails += run_case(128, 256, 256, EPI_BF16, ACT_NONE, 256, false, num_sms);
  fails += run_case(128, 256, 1152, EPI_BF16, ACT_NONE, 256, false, num_sms);
  // tails in M, N, K
  fails += run_case(50, 512, 768, EPI_BF16, ACT_QUICKGELU, 0, false, num_sms);
  fails += run_case(300, 1152, 1152, EPI_BF16, ACT_GELU_TANH, 0, false, num_sms);
  fails += run_case(389, 4304, 1152, EPI_BF16, ACT_GELU_ERF, 0, false, num_sms);
  fails += run_case(389, 1152, 4304, EPI_RESID, ACT_NONE, 0, false, num_sms);
  fails += run_case(343, 768, 592, EPI_F32, ACT_NONE, 0, false, num_sms);
  fails += run_case(77 * 5, 1024, 1024, EPI_RESID, ACT_NONE, 0, false, num_sms);
  // CTA-pair (cta_group::2) variants: one pair tile, tails, several tiles per pair
  fails += run_case(256, 256, 64, EPI_BF16, ACT_NONE, 256, false, num_sms, 2);
  fails += run_case(256, 128, 128, EPI_BF16, ACT_NONE, 128, false, num_sms, 2);
  fails += run_case(256, 192, 1152, EPI_RESID, ACT_NONE, 192, false, num_sms, 2);
  fails += run_case(389, 4304, 1152, EPI_BF16, ACT_GELU_TANH, 0, false, num_sms, 2);
  fails += run_case(1000, 1152, 4304, EPI_RESID, ACT_NONE, 0, false, num_sms, 2);
  fails += run_case(343, 768, 592, EPI_F32, ACT_NONE, 0, false, num_sms, 2);
  if (!quick) {
    // multi-tile persistent scheduling + perf (SO400M layer shapes at a 128-image micro-batch, M = 73728)
    const int M = 128 * 576;
    fails += run_case(M, 3456, 1152, EPI_BF16, ACT_NONE, 0, true, num_sms);
    fails += run_case(M, 3456, 1152, EPI_BF16, ACT_NONE, 256, true, num_sms);
    fails += run_case(M, 1152, 1152, EPI_RESID, ACT_NONE, 0, true, num_sms);
    fails += run_case(M, 4304, 1152, EPI_BF16, ACT_GELU_TANH, 0, true, num_sms);
    fails += run_case(M, 4304, 1152, EPI_BF16, ACT_GELU_TANH, 256, true, num_sms);
    fails += run_case(M, 1152, 4304, EPI_RESID, ACT_NONE, 0, true, num_sms);
    fails += run_case(M, 1152, 4304, EPI_RESID, ACT_NONE, 128, true, num_sms);
    fails += run_case(8192, 8192, 8192, EPI_BF16, ACT_NONE, 256, true, num_sms);
    fails += run_case(M, 3456, 1152, EPI_BF16, ACT_NONE, 256, true, num_sms, 2);
    fails += run_case(M, 1152, 1152, EPI_RESID, ACT_NONE, 0, true, num_sms, 2);
    fails += run_case(M, 4304, 1152, EPI_BF16, ACT_GELU_TANH, 256, true, num_sms, 2);
    fails += run_case(M, 1152, 4304, EPI_RESID, ACT_NONE, 0, true, num_sms, 2);
    fails += run_case(8192, 8192, 8192, EPI_BF16, ACT_NONE, 256, true, num_sms, 2);
  }
  printf("%s (%d failing cases)\n", fails ? "GEMM TEST FAILED" : "GEMM TEST PASSED", fails);
  return fails ? 1 : 0;
}
