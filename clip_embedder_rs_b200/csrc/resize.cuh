// GPU resize (reference src/vision.rs:164-198).  See resize.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace clipb200 {

struct ResizeAxis {  // host-side coefficient table of one axis
  std::vector<int32_t> start, size;
  std::vector<int16_t> w;  // [out][window]
  int window = 0, precision = 0;
};
struct ResizePlanDev {  // device-resident plan for one (source size, target size, filter, crop mode)
  bool nearest = false;
  int *xstart = nullptr, *xsize = nullptr, *ystart = nullptr, *ysize = nullptr;
  int16_t *xw = nullptr, *yw = nullptr;
  int xwindow = 0, ywindow = 0, xprecision = 0, yprecision = 0;
  int y_first = 0, rows = 0;
  double left = 0, top = 0, sx = 1, sy = 1;
};

void resize_crop_box(int width, int height, int size, bool squash, double* left, double* top, double* cw, double* ch);
ResizeAxis make_resize_axis(int in_size, double in0, double in1, int out_size, int interpolation /*0 cubic, 1 linear*/);
// d_src [H,W,3] u8 -> d_dst [S,S,3] u8; d_tmp holds plan.rows * S * 3 bytes
cudaError_t launch_resize(const uint8_t* d_src, int W, int H, int S, const ResizePlanDev& plan, uint8_t* d_tmp, uint8_t* d_dst,
                          cudaStream_t st);

}  // namespace clipb200
