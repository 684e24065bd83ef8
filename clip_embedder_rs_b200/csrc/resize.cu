// GPU image resize replacing `resize_with_fast_image_resize` (reference src/vision.rs:164-198): antialiased two-pass
// convolution (CatmullRom for "bicubic", triangle for "bilinear") or nearest, with the reference's f64 centre-crop
// box for every resize_mode except "squash".  Arithmetic follows fast_image_resize 6.0.0's U8x3 path as recalled
// (Pillow-SIMD scheme): f64 weights normalised per output pixel, converted to i16 at the largest precision that
// keeps the biggest weight below 2^15, integer accumulation, u8 intermediate between the horizontal and the vertical
// pass.  The coefficient tables are built on the host (tiny, cached per axis by the engine) and travel with each group of
// images; a whole group is resized by two launches (blockIdx.z = image); the passes are HBM-bound.
#include "resize.cuh"

#include <math.h>

#include <algorithm>

namespace clipb200 {

static double catmull_rom(double x) {
  const double a = -0.5;
  x = fabs(x);
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0;
  if (x < 2.0) return (((x - 5.0) * x + 8.0) * x - 4.0) * a;
  return 0.0;
}
static double triangle(double x) {
  x = fabs(x);
  return x < 1.0 ? 1.0 - x : 0.0;
}

void resize_crop_box(int width, int height, int size, bool squash, double* left, double* top, double* cw, double* ch) {
  if (squash) {
    *left = 0.0; *top = 0.0; *cw = width; *ch = height;
    return;
  }
  // vision.rs:184-192, same f64 operation order
  const double scale = static_cast<double>(size) / static_cast<double>(std::min(width, height));
  *cw = static_cast<double>(size) / scale;
  *ch = static_cast<double>(size) / scale;
  *left = (static_cast<double>(width) - *cw) / 2.0;
  *top = (static_cast<double>(height) - *ch) / 2.0;
}

ResizeAxis make_resize_axis(int in_size, double in0, double in1, int out_size, int interpolation) {
  ResizeAxis ax;
  const bool cubic = interpolation == 0;
  const double support = cubic ? 2.0 : 1.0;
  const double scale = (in1 - in0) / static_cast<double>(out_size);
  const double filter_scale = std::max(scale, 1.0);
  const double radius = support * filter_scale;
  ax.window = static_cast<int>(ceil(radius)) * 2 + 1;
  ax.start.assign(out_size, 0);
  ax.size.assign(out_size, 0);
  std::vector<double> w(static_cast<size_t>(out_size) * ax.window, 0.0);
  double max_w = 0.0;
  for (int o = 0; o < out_size; ++o) {
    const double centre = in0 + (o + 0.5) * scale;
    const int x_min = static_cast<int>(std::max(floor(centre - radius), 0.0));
    const int x_max = static_cast<int>(std::min(ceil(centre + radius), static_cast<double>(in_size)));
    const double c = centre - 0.5;
    double total = 0.0;
    double* row = &w[static_cast<size_t>(o) * ax.window];
    for (int x = x_min; x < x_max; ++x) {
      const double v = cubic ? catmull_rom((x - c) / filter_scale) : triangle((x - c) / filter_scale);
      row[x - x_min] = v;
      total += v;
    }
    if (total != 0.0)
      for (int i = 0; i < x_max - x_min; ++i) row[i] /= total;
    ax.start[o] = x_min;
    ax.size[o] = x_max - x_min;
  }
  for (double v : w) max_w = std::max(max_w, v);
  int precision = 0;
  for (int cur = 0; cur < 16; ++cur) {
    precision = cur;
    if (static_cast<int>(llround(max_w * static_cast<double>(1 << (cur + 1)))) >= (1 << 15)) break;
  }
  ax.precision = precision;
  ax.w.resize(w.size());
  for (size_t i = 0; i < w.size(); ++i) ax.w[i] = static_cast<int16_t>(llround(w[i] * static_cast<double>(1 << precision)));
  return ax;
}

// ---- batched variants: blockIdx.z = image of the group ------------------------------------------------------------
__global__ void __launch_bounds__(256)
resize_h_batched_kernel(const uint8_t* __restrict__ src, const ResizeJob* __restrict__ jobs,
                        const int32_t* __restrict__ arena, int S, uint8_t* __restrict__ tmp) {
  const ResizeJob& j = jobs[blockIdx.z];
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (j.mode != 0 || ox >= S || r >= j.rows) return;
  const int x0 = arena[j.xstart + ox], n = arena[j.xsize + ox];
  // staged row r = columns [x_first, x_first + pitch) of source row y_first + r
  const uint8_t* row = src + j.src_off + (static_cast<long long>(r) * j.pitch + (x0 - j.x_first)) * 3;
  const int16_t* ww = reinterpret_cast<const int16_t*>(arena + j.xw) + static_cast<long long>(ox) * j.xwindow;
  const int half = j.xprecision > 0 ? (1 << (j.xprecision - 1)) : 0;
  int a0 = half, a1 = half, a2 = half;
  for (int i = 0; i < n; ++i) {
    const int c = ww[i];
    a0 += row[3 * i] * c; a1 += row[3 * i + 1] * c; a2 += row[3 * i + 2] * c;
  }
  uint8_t* o = tmp + j.tmp_off + (static_cast<long long>(r) * S + ox) * 3;
  o[0] = static_cast<uint8_t>(min(max(a0 >> j.xprecision, 0), 255));
  o[1] = static_cast<uint8_t>(min(max(a1 >> j.xprecision, 0), 255));
  o[2] = static_cast<uint8_t>(min(max(a2 >> j.xprecision, 0), 255));
}
__global__ void __launch_bounds__(256)
resize_v_batched_kernel(const uint8_t* __restrict__ src, const uint8_t* __restrict__ tmp,
                        const ResizeJob* __restrict__ jobs, const int32_t* __restrict__ arena, int S,
                        uint8_t* __restrict__ dst) {
  const ResizeJob& j = jobs[blockIdx.z];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over S*3 bytes of one output row
  const int oy = blockIdx.y;
  if (idx >= S * 3) return;
  uint8_t* out = dst + j.dst_off + static_cast<long long>(oy) * S * 3 + idx;
  if (j.mode == 2) {  // already at the model resolution: the convolution is the identity
    *out = src[j.src_off + static_cast<long long>(oy) * S * 3 + idx];
    return;
  }
  if (j.mode == 1) {  // ResizeAlg::Nearest (vision.rs:179)
    const int ox = idx / 3, c = idx - ox * 3;
    const int x = min(static_cast<int>(floor(j.left + (ox + 0.5) * j.sx)), j.W - 1);
    const int y = min(static_cast<int>(floor(j.top + (oy + 0.5) * j.sy)), j.H - 1);
    *out = src[j.src_off + (static_cast<long long>(y) * j.W + x) * 3 + c];
    return;
  }
  const int y0 = arena[j.ystart + oy], n = arena[j.ysize + oy];
  const int16_t* ww = reinterpret_cast<const int16_t*>(arena + j.yw) + static_cast<long long>(oy) * j.ywindow;
  const uint8_t* col = tmp + j.tmp_off + static_cast<long long>(y0 - j.y_first) * S * 3 + idx;
  int acc = j.yprecision > 0 ? (1 << (j.yprecision - 1)) : 0;
  for (int i = 0; i < n; ++i) acc += col[static_cast<long long>(i) * S * 3] * static_cast<int>(ww[i]);
  *out = static_cast<uint8_t>(min(max(acc >> j.yprecision, 0), 255));
}

cudaError_t launch_resize_batched(const uint8_t* d_src, const ResizeJob* d_jobs, const int32_t* d_arena, int n, int S,
                                  int max_rows, uint8_t* d_tmp, uint8_t* d_dst, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (n > 65535 || max_rows > 65535) return cudaErrorInvalidValue;
  if (max_rows > 0) {
    resize_h_batched_kernel<<<dim3((S + 255) / 256, max_rows, n), 256, 0, st>>>(d_src, d_jobs, d_arena, S, d_tmp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  resize_v_batched_kernel<<<dim3((S * 3 + 255) / 256, S, n), 256, 0, st>>>(d_src, d_tmp, d_jobs, d_arena, S, d_dst);
  return cudaGetLastError();
}

}  // namespace clipb200
