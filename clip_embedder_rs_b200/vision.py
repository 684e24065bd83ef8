"""`VisionEmbedder` — mirrors `/root/reference/src/vision.rs:21-259`.

Images are `numpy.uint8` arrays of shape [H, W, 3] (RGB), or anything with `.convert("RGB")` (PIL), standing in
for `image::DynamicImage`.  `embed_images` hands the packed RGB8 batch to the engine, which normalises on the GPU
(vision.rs:235-259) and runs the tower; nothing is computed on the CPU."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

from . import _native, error, model_manager
from .config import ModelConfig, OpenClipConfig
from .onnx import OnnxSession

_INTERP = {"bicubic": 0, "bilinear": 1}
_RESIZE = {"squash": 1}


def _to_rgb8(image) -> np.ndarray:
    """`image.to_rgb8()` (vision.rs:171)."""
    if isinstance(image, np.ndarray):
        a = image
    elif hasattr(image, "convert"):
        a = np.asarray(image.convert("RGB"))
    else:
        raise error.Image(f"Image error: unsupported image type {type(image)!r}")
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        raise error.Image(f"Image error: expected uint8 [H,W,3], got {a.dtype} {a.shape}")
    return a


class _Builder:
    """Stands in for the bon-generated builders (`finish_fn = build`, vision.rs:29-84)."""

    def __init__(self, cls, model_dir: Path):
        self._cls, self._dir, self._eps, self._kw = cls, Path(model_dir), None, {}

    def with_execution_providers(self, eps):
        self._eps = eps
        return self

    def maybe_with_execution_providers(self, eps):
        self._eps = eps
        return self

    def device(self, index: int):
        self._kw["device"] = int(index)
        return self

    def devices(self, indices: Sequence[int]):
        """In-process multi-GPU pool: one replica per listed CUDA device (`[]` = all visible), batches split row-wise
        (what N `duplicate()`s of vision.rs:87-91 and a caller-side split would do)."""
        self._kw["devices"] = [int(i) for i in indices]
        return self

    def micro_batch(self, n: int):
        self._kw["micro_batch"] = int(n)
        return self

    def profile(self, on: bool = True):
        self._kw["profile"] = bool(on)
        return self

    def build(self):
        return self._cls._load(self._dir, self._eps, **self._kw)


class _IdBuilder(_Builder):
    def __init__(self, cls, model_id: str):
        super().__init__(cls, Path("."))
        self._id, self._base = model_id, None

    def base_folder(self, p):
        self._base = Path(p)
        return self

    def build(self):
        base = self._base if self._base is not None else model_manager.get_default_base_folder()
        return self._cls._load(base / self._id, self._eps, **self._kw)


class _HfBuilder(_Builder):
    def __init__(self, cls, model_id: str):
        super().__init__(cls, Path("."))
        self._id = model_id

    def build(self):
        return self._cls._load(model_manager.get_hf_model(self._id), self._eps, **self._kw)


class VisionEmbedder:
    # pub fields of the reference struct (vision.rs:21-27)
    session: OnnxSession
    config: OpenClipConfig
    model_config: ModelConfig
    input_name: str
    model_dir: Path

    @classmethod
    def from_hf(cls, model_id: str) -> _Builder:  # vision.rs:32-42
        return _HfBuilder(cls, model_id)

    @classmethod
    def from_local_id(cls, model_id: str) -> _IdBuilder:  # vision.rs:45-55
        return _IdBuilder(cls, model_id)

    @classmethod
    def from_local_dir(cls, model_dir) -> _Builder:  # vision.rs:58-84
        return _Builder(cls, Path(model_dir))

    @classmethod
    def _load(cls, model_dir: Path, execution_providers=None, device: int = 0, micro_batch: int = 0,
              profile: bool = False, devices=None) -> "VisionEmbedder":
        model_manager.verify_model_dir(model_dir)
        self = cls.__new__(cls)
        self.session = OnnxSession(model_dir / "visual.onnx", execution_providers, device, micro_batch, profile, devices)
        self.config = OpenClipConfig.from_file(model_dir / "open_clip_config.json")
        self.model_config = ModelConfig.from_file(model_dir / "model_config.json")
        name = self.session.find_input(["pixel_values", "input"])
        if name is None:
            raise error.Config("Could not find vision input node")
        self.input_name = name
        self.model_dir = Path(model_dir)
        self._kw = dict(device=device, micro_batch=micro_batch, profile=profile, devices=devices)
        pc = self.config.preprocess_cfg
        self._pp = _native.Preproc((pc.mean[0], pc.mean[1], pc.mean[2]), (pc.std[0], pc.std[1], pc.std[2]),
                                   _INTERP.get(pc.interpolation, 2), _RESIZE.get(pc.resize_mode, 0))
        return self

    def duplicate(self) -> "VisionEmbedder":  # vision.rs:87-91
        return type(self)._load(self.model_dir, self.session.execution_providers, **self._kw)

    # ---------------------------------------------------------------------------------------------- hot path
    def _rgb_list(self, images: Sequence):
        if len(images) == 0:
            raise error.Inference("Empty batch")  # vision.rs:121-123
        if isinstance(images, np.ndarray) and images.ndim == 4:
            if images.dtype != np.uint8 or images.shape[3] != 3:
                raise error.Image(f"Image error: expected uint8 [B,H,W,3], got {images.dtype} {images.shape}")
            # the engine reads packed RGB through raw pointers: a strided view (rgba[..., :3], a flip) must be copied
            return np.ascontiguousarray(images)
        return [np.ascontiguousarray(_to_rgb8(im)) for im in images]

    def _pack(self, images: Sequence):
        """Returns a packed [B,S,S,3] array when every image is already at the model resolution, else None."""
        size = self.config.model_cfg.vision_cfg.image_size
        arrs = self._rgb_list(images)
        if all(a.shape[0] == size and a.shape[1] == size for a in arrs):
            return np.ascontiguousarray(arrs if isinstance(arrs, np.ndarray) else np.stack(arrs, axis=0)), arrs
        return None, arrs

    @staticmethod
    def _pointer_arrays(arrs):
        n = len(arrs)
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        ws = np.asarray([a.shape[1] for a in arrs], dtype=np.int32)
        hs = np.asarray([a.shape[0] for a in arrs], dtype=np.int32)
        return ptrs, ws, hs

    def embed_image(self, image) -> np.ndarray:  # vision.rs:94-98
        return self.embed_images([image]).reshape(-1)

    def embed_images(self, images: Sequence) -> np.ndarray:  # vision.rs:102-117
        batch, arrs = self._pack(images)
        n = len(arrs)
        out = np.empty((n, self.session.embed_dim), dtype=np.float32)
        if batch is not None:
            size = batch.shape[1]
            self.session.run_rgb8(batch.ctypes.data, n, size, size, self._pp, out.ctypes.data)
        else:  # arbitrary sizes: resize (vision.rs:164-198) on the GPU
            ptrs, ws, hs = self._pointer_arrays(arrs)
            self.session.run_rgb8_var(ptrs, ws.ctypes.data, hs.ctypes.data, n, self._pp, out.ctypes.data)
        return out

    def resize(self, image) -> np.ndarray:
        """`resize_with_fast_image_resize` (vision.rs:164-198): any size in, [S,S,3] uint8 out (computed on the GPU)."""
        a = np.ascontiguousarray(_to_rgb8(image))
        size = self.config.model_cfg.vision_cfg.image_size
        out = np.empty((size, size, 3), dtype=np.uint8)
        with self.session._lock:
            self.session.check(_native.lib.clipb200_resize_rgb8(self.session.handle, a.ctypes.data, a.shape[1], a.shape[0],
                                                               self._pp, out.ctypes.data))
        return out

    def embed_pixel_values(self, pixel_values: np.ndarray) -> np.ndarray:
        """What `session.run(inputs![pixel_values])` does in the reference (vision.rs:105-113): f32 [B,3,S,S] in."""
        pv = np.ascontiguousarray(pixel_values, dtype=np.float32)
        size = self.config.model_cfg.vision_cfg.image_size
        if pv.ndim != 4 or pv.shape[1] != 3 or pv.shape[2] != size or pv.shape[3] != size:
            raise error.Shape(f"Shape error: pixel_values must be [B,3,{size},{size}], got {pv.shape}")
        if pv.shape[0] == 0:
            raise error.Inference("Empty batch")
        out = np.empty((pv.shape[0], self.session.embed_dim), dtype=np.float32)
        with self.session._lock:
            self.session.check(_native.lib.clipb200_vision_embed_f32(self.session.handle, pv.ctypes.data,
                                                                     pv.shape[0], out.ctypes.data))
        return out

    def preprocess_batch(self, images: Sequence) -> np.ndarray:  # vision.rs:120-135
        batch, arrs = self._pack(images)
        if batch is None:
            batch = np.stack([self.resize(a) for a in arrs], axis=0)
        n, size = batch.shape[0], batch.shape[1]
        out = np.empty((n, 3, size, size), dtype=np.float32)
        with self.session._lock:
            self.session.check(_native.lib.clipb200_preprocess_rgb8(
                self.session.handle, batch.ctypes.data, n, size, size, self._pp, out.ctypes.data))
        return out

    def preprocess(self, image) -> np.ndarray:  # vision.rs:138-140
        return self.preprocess_batch([image])
