"""`OnnxSession` — mirrors `/root/reference/src/onnx.rs:8-47`, with the `ort::Session` replaced by a handle to the
B200 engine (`clipb200_engine*`).  `execution_providers` is kept for API compatibility; the only provider this build
has is the sm_100a engine, optionally addressed by a CUDA device index."""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import List, Optional, Sequence

from . import _native, error


def _raise_engine_error(code: int) -> None:
    raise error.Ort(_native.last_error(), code)


def inspect_onnx(path) -> dict:
    """Parse-only description of a model file (no GPU): inputs, outputs, opset, metadata and — for files that carry
    an executable graph — how the graph recogniser bound every parameter (`["graph"]["bindings"]`)."""
    import json

    cap = 1 << 20
    while True:
        buf = C.create_string_buffer(cap)
        rc = _native.lib.clipb200_onnx_inspect(os.fspath(path).encode(), buf, cap)
        if rc == _native.OK:
            return json.loads(buf.value.decode())
        if "too small" in _native.last_error() and cap < (1 << 28):
            cap *= 4
            continue
        _raise_engine_error(rc)


def read_onnx_tensor(path, name: str):
    """fp32 value of one parameter exactly as the engine would bind it (canonical open_clip / timm name, `[out, in]`
    layout for Linear weights), without touching a GPU."""
    import numpy as np

    dims = (C.c_int64 * 8)()
    rank = C.c_int(0)
    p = os.fspath(path).encode()
    rc = _native.lib.clipb200_onnx_read_tensor(p, name.encode(), None, 0, dims, C.byref(rank))
    if rc != _native.OK:
        _raise_engine_error(rc)
    shape = tuple(int(dims[i]) for i in range(rank.value))
    out = np.empty(shape, dtype=np.float32)
    rc = _native.lib.clipb200_onnx_read_tensor(p, name.encode(), out.ctypes.data_as(C.c_void_p), out.size, dims, C.byref(rank))
    if rc != _native.OK:
        _raise_engine_error(rc)
    return out


class OnnxSession:
    def __init__(self, path, execution_providers: Optional[Sequence] = None, device: int = 0,
                 micro_batch: int = 0, profile: bool = False, devices: Optional[Sequence[int]] = None):
        """onnx.rs:14-29 (`OnnxSession::new`).  `devices` (a list of CUDA indices, `[]` = all visible GPUs) makes this
        session an in-process pool of replicas — `duplicate()` (vision.rs:87-91) for the GPUs of one box — with the
        batch of every run split row-wise over them (`clipb200_pool_*`)."""
        self.execution_providers: List = list(execution_providers or [])
        self.device = int(device)
        self.devices: Optional[List[int]] = None if devices is None else [int(d) for d in devices]
        self.path = os.fspath(path)
        self._lock = threading.RLock()  # the reference serialises runs with RwLock::write (vision.rs:107)
        opts = _native.Opts(micro_batch=int(micro_batch), profile=1 if profile else 0)
        handle = C.c_void_p()
        self._h = None
        self._pool = None
        if self.devices is None:
            rc = _native.lib.clipb200_engine_create(self.path.encode(), self.device, C.byref(opts), C.byref(handle))
            if rc != _native.OK:
                _raise_engine_error(rc)
            self._h = handle
        else:
            arr = (C.c_int32 * max(1, len(self.devices)))(*self.devices)
            rc = _native.lib.clipb200_pool_create(self.path.encode(), arr, len(self.devices), C.byref(opts),
                                                  C.byref(handle))
            if rc != _native.OK:
                _raise_engine_error(rc)
            self._pool = handle
            self.devices = [int(_native.lib.clipb200_pool_device(handle, i))
                            for i in range(_native.lib.clipb200_pool_size(handle))]
            self.device = self.devices[0]

    @property
    def is_pool(self) -> bool:
        return self._pool is not None

    # -- introspection (onnx.rs:32-46)
    def input_names(self) -> List[str]:
        if self._pool is not None:
            n = _native.lib.clipb200_pool_num_inputs(self._pool)
            return [_native.lib.clipb200_pool_input_name(self._pool, i).decode() for i in range(n)]
        n = _native.lib.clipb200_engine_num_inputs(self._h)
        return [_native.lib.clipb200_engine_input_name(self._h, i).decode() for i in range(n)]

    def has_input(self, name: str) -> bool:
        return name in self.input_names()

    def find_input(self, possibilities: Sequence[str]) -> Optional[str]:
        names = self.input_names()
        for p in possibilities:
            if p in names:
                return p
        return None

    @property
    def handle(self) -> C.c_void_p:
        if self._h is None:
            raise error.Config("this session is a multi-GPU pool: it has no single engine handle")
        return self._h

    def _which(self, engine_value, pool_fn):
        return int(pool_fn(self._pool)) if self._pool is not None else int(engine_value(self._h))

    @property
    def embed_dim(self) -> int:
        return self._which(_native.lib.clipb200_engine_embed_dim, _native.lib.clipb200_pool_embed_dim)

    @property
    def image_size(self) -> int:
        return self._which(_native.lib.clipb200_engine_image_size, _native.lib.clipb200_pool_image_size)

    @property
    def context_length(self) -> int:
        return self._which(_native.lib.clipb200_engine_context_length, _native.lib.clipb200_pool_context_length)

    # -- `session.run` (vision.rs:105-113, text.rs:153-166): host buffers in, host buffers out
    def run_rgb8(self, hwc_ptr, n: int, width: int, height: int, pp, out_ptr) -> None:
        with self._lock:
            if self._pool is not None:
                self.check(_native.lib.clipb200_pool_vision_embed_rgb8(self._pool, hwc_ptr, n, width, height, pp, out_ptr))
            else:
                self.check(_native.lib.clipb200_vision_embed_rgb8(self._h, hwc_ptr, n, width, height, pp, out_ptr))

    def run_rgb8_var(self, ptrs, widths_ptr, heights_ptr, n: int, pp, out_ptr) -> None:
        with self._lock:
            if self._pool is not None:
                self.check(_native.lib.clipb200_pool_vision_embed_rgb8_var(self._pool, ptrs, widths_ptr, heights_ptr, n, pp,
                                                                           out_ptr))
            else:
                self.check(_native.lib.clipb200_vision_embed_rgb8_var(self._h, ptrs, widths_ptr, heights_ptr, n, pp, out_ptr))

    def run_ids(self, ids_ptr, mask_ptr, n: int, ctx: int, out_ptr) -> None:
        with self._lock:
            if self._pool is not None:
                self.check(_native.lib.clipb200_pool_text_embed(self._pool, ids_ptr, mask_ptr, n, ctx, out_ptr))
            else:
                self.check(_native.lib.clipb200_text_embed(self._h, ids_ptr, mask_ptr, n, ctx, out_ptr))

    @property
    def weight_bytes(self) -> int:
        return int(_native.lib.clipb200_engine_weight_bytes(self.handle))

    @property
    def launch_count(self) -> int:
        if self._pool is not None:
            return int(_native.lib.clipb200_pool_launch_count(self._pool))
        return int(_native.lib.clipb200_engine_launch_count(self._h))

    def check(self, rc: int) -> None:
        if rc != _native.OK:
            _raise_engine_error(rc)

    def synchronize(self) -> None:
        self.check(_native.lib.clipb200_engine_synchronize(self.handle))

    def profile(self, reset: bool = True) -> dict:
        p = _native.Profile()
        self.check(_native.lib.clipb200_engine_profile(self.handle, C.byref(p), 1 if reset else 0))
        return {"ms": {n: p.ms[i] for i, n in enumerate(_native.PROF_NAMES)},
                "launches": {n: int(p.launches[i]) for i, n in enumerate(_native.PROF_NAMES)},
                "gemm_flops": p.gemm_flops, "conv_bytes": p.conv_bytes}

    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            _native.lib.clipb200_engine_destroy(h)
        p, self._pool = getattr(self, "_pool", None), None
        if p:
            _native.lib.clipb200_pool_destroy(p)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
