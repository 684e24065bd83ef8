"""CPU ORACLE for the image resize of `resize_with_fast_image_resize` (`/root/reference/src/vision.rs:164-198`).
Test infrastructure only.  PARITY UNPINNED for the third-party part:

The reference calls `fast_image_resize` 6.0.0 (`Cargo.lock:792-793`, not vendored, no Rust toolchain here):
`Resizer::resize(src U8x3, dst U8x3, ResizeOptions{ResizeAlg::Convolution(CatmullRom | Bilinear) | Nearest, crop})`.
What is restated here is that crate's published algorithm as recalled (it is the Pillow-SIMD scheme):

  * per output coordinate a window of source pixels [x_min, x_max) around centre = in0 + (out + 0.5) * scale with
    radius support * max(scale, 1) (antialiasing: the kernel widens when downscaling), weights
    kernel((x - centre + 0.5) / max(scale, 1)) normalised to sum 1 in f64;
  * weights converted to i16 with the largest precision p such that round(max_w * 2^(p+1)) < 2^15;
  * two passes, horizontal then vertical, u8 intermediate; each output = clip8((2^(p-1) + sum px * w_i16) >> p);
  * the crop box (f64 left/top/width/height) of `vision.rs:184-192` supplies in0 / in1 per axis;
  * `ResizeAlg::Nearest`: src = floor(in0 + (out + 0.5) * scale).

`tests/test_resize_cpu.py` cross-checks it against Pillow (same windows and kernels, 22-bit instead of <=15-bit
coefficients): at most 1 LSB apart on structured noise and — `test_reference_photos_against_pillow`, which opens
`/root/reference/assets/img/*.jpg` — on the reference's seven example photos at 224 / 256 / 384 px, centre-cropped
and squashed; `tests/test_resize_gpu.py::test_reference_photos_bit_exact` holds the GPU kernels to this oracle byte
for byte on the same photos.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np


def catmull_rom(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    if x < 2.0:
        return (((x - 5.0) * x + 8.0) * x - 4.0) * a
    return 0.0


def bilinear(x: float) -> float:
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


FILTERS = {"bicubic": (catmull_rom, 2.0), "bilinear": (bilinear, 1.0)}


def crop_box(width: int, height: int, size: int, resize_mode: str) -> Tuple[float, float, float, float]:
    """vision.rs:184-192 (f64 arithmetic, same operation order)."""
    if resize_mode == "squash":
        return 0.0, 0.0, float(width), float(height)
    scale = float(size) / float(min(width, height))
    crop_w = float(size) / scale
    crop_h = float(size) / scale
    return (float(width) - crop_w) / 2.0, (float(height) - crop_h) / 2.0, crop_w, crop_h


def precompute_coefficients(in_size: int, in0: float, in1: float, out_size: int, kernel, support: float):
    scale = (in1 - in0) / float(out_size)
    filter_scale = max(scale, 1.0)
    radius = support * filter_scale
    window = int(math.ceil(radius)) * 2 + 1
    starts = np.zeros(out_size, dtype=np.int32)
    sizes = np.zeros(out_size, dtype=np.int32)
    weights = np.zeros((out_size, window), dtype=np.float64)
    for o in range(out_size):
        centre = in0 + (o + 0.5) * scale
        x_min = int(max(math.floor(centre - radius), 0.0))
        x_max = int(min(math.ceil(centre + radius), float(in_size)))
        c = centre - 0.5
        ws = [kernel((x - c) / filter_scale) for x in range(x_min, x_max)]
        total = sum(ws)
        if total != 0.0:
            ws = [w / total for w in ws]
        starts[o], sizes[o] = x_min, x_max - x_min
        weights[o, :len(ws)] = ws
    return starts, sizes, weights


def _round_half_away(x):
    """Rust `f64::round`: halves round away from zero (numpy / Python round half to even)."""
    x = np.asarray(x, dtype=np.float64)
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def normalise_i16(weights: np.ndarray) -> Tuple[np.ndarray, int]:
    max_w = float(weights.max()) if weights.size else 0.0
    precision = 0
    for cur in range(16):
        precision = cur
        if int(_round_half_away(max_w * (1 << (cur + 1)))) >= (1 << 15):
            break
    return _round_half_away(weights * (1 << precision)).astype(np.int16), precision


def _convolve_rows(src: np.ndarray, starts, sizes, w16, precision) -> np.ndarray:
    """src [N, in, 3] u8 -> [N, out, 3] u8 along axis 1."""
    out = np.empty((src.shape[0], len(starts), 3), dtype=np.uint8)
    half = 1 << (precision - 1) if precision > 0 else 0
    s32 = src.astype(np.int32)
    for o in range(len(starts)):
        a, n = int(starts[o]), int(sizes[o])
        acc = (s32[:, a:a + n, :] * w16[o, :n].astype(np.int32)[None, :, None]).sum(axis=1) + half
        out[:, o, :] = np.clip(acc >> precision, 0, 255).astype(np.uint8)
    return out


def resize_rgb8(img: np.ndarray, size: int, interpolation: str = "bicubic", resize_mode: str = "shortest") -> np.ndarray:
    """img u8 [H, W, 3] -> u8 [size, size, 3], following vision.rs:164-198."""
    h, w = img.shape[0], img.shape[1]
    left, top, cw, ch = crop_box(w, h, size, resize_mode)
    if interpolation not in FILTERS:  # ResizeAlg::Nearest (vision.rs:179)
        sx, sy = cw / size, ch / size
        xs = np.minimum(np.floor(left + (np.arange(size) + 0.5) * sx).astype(np.int64), w - 1)
        ys = np.minimum(np.floor(top + (np.arange(size) + 0.5) * sy).astype(np.int64), h - 1)
        return img[ys][:, xs]
    kernel, support = FILTERS[interpolation]
    xs, xn, xw = precompute_coefficients(w, left, left + cw, size, kernel, support)
    ys, yn, yw = precompute_coefficients(h, top, top + ch, size, kernel, support)
    xw16, xp = normalise_i16(xw)
    yw16, yp = normalise_i16(yw)
    y_first = int(ys.min())
    y_last = int((ys + yn).max())
    tmp = _convolve_rows(img[y_first:y_last], xs, xn, xw16, xp)                  # horizontal pass, u8 intermediate
    out = _convolve_rows(tmp.transpose(1, 0, 2), ys - y_first, yn, yw16, yp)     # vertical pass
    return np.ascontiguousarray(out.transpose(1, 0, 2))
