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
// One image of a batched resize (clipb200_vision_embed_rgb8_var): offsets into the group's source staging buffer, its
// coefficient arena (int32 words: per axis `start[S]`, `size[S]`, then `w[S][window]` as int16) and its intermediate.
struct ResizeJob {
  long long src_off, tmp_off, dst_off;  // bytes
  int W, H;
  int mode;                             // 0 two-pass convolution, 1 nearest, 2 copy (already S x S)
  int y_first, rows;                    // source rows [y_first, y_first + rows) are staged at src_off (mode 0)
  int x_first, pitch;                   // ... and of each of them the columns [x_first, x_first + pitch) (mode 0; else 0, W)
  int xstart, xsize, xw, ystart, ysize, yw;  // word offsets into the arena
  int xwindow, ywindow, xprecision, yprecision;
  double left, top, sx, sy;             // nearest only
};
// Resizes `n` images with two launches (horizontal pass into d_tmp, vertical pass / nearest / copy into d_dst).
// `max_rows` = the largest job.rows of the group (grid height of the horizontal pass).
cudaError_t launch_resize_batched(const uint8_t* d_src, const ResizeJob* d_jobs, const int32_t* d_arena, int n, int S,
                                  int max_rows, uint8_t* d_tmp, uint8_t* d_dst, cudaStream_t st);

void resize_crop_box(int width, int height, int size, bool squash, double* left, double* top, double* cw, double* ch);
ResizeAxis make_resize_axis(int in_size, double in0, double in1, int out_size, int interpolation /*0 cubic, 1 linear*/);

}  // namespace clipb200
