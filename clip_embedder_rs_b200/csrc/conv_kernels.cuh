// Convolution-stage kernels of the FastViT / MobileCLIP2 hybrid trunk.  See conv_kernels.cu.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace clipb200 {

// image (uint8 HWC through the [3][256] LUT, or f32 NCHW) -> bf16 NHWC [n, S/2, S/2, Cout], GELU fused
cudaError_t launch_stem_conv3x3_s2(const uint8_t* img_u8, const float* img_f32, const float* lut, int n, int S, int Cout,
                                   const float* w27 /*[27][Cout]*/, const float* bias, __nv_bfloat16* out, cudaStream_t st);
// depthwise KxK conv, NHWC; weight [K*K][Cin*mult]; in/out fp32 or bf16
cudaError_t launch_dwconv(const void* in, bool in_bf16, int n, int H, int W, int Cin, int K, int stride, int mult,
                          const float* w, const float* bias, bool gelu, void* out, bool out_bf16, cudaStream_t st);
// stride-1, multiplier-1 fp32 variants on the TMA-staged kernel (dwconv_sm100.cu); launch_dwconv routes to it
cudaError_t dwconv_tma_configure_device();
bool dwconv_tma_supported(bool in_bf16, int C, int K, int stride, int mult, bool gelu);
cudaError_t launch_dwconv_tma(const float* in, int n, int H, int W, int C, int K, const float* w, const float* bias,
                              void* out, bool out_bf16, cudaStream_t st);
bool launch_dwconv_tma_gen(const void* in, bool in_bf16, int n, int H, int W, int C, int K, int stride, int mult,
                           const float* w, const float* bias, bool gelu, void* out, bool out_bf16, cudaStream_t st,
                           cudaError_t* err);
cudaError_t launch_gap(const float* x, int n, int P, int C, float* out, cudaStream_t st);
cudaError_t launch_se_mlp(const float* s, int n, int C, int R, const float* w1, const float* b1, const float* w2,
                          const float* b2, float* gate, cudaStream_t st);
cudaError_t launch_scale_act(const float* x, const float* gate_or_null, int n, int P, int C, bool gelu, void* out,
                             bool out_bf16, cudaStream_t st);

}  // namespace clipb200
