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
                 micro_batch: int = 0, profile: bool = False):
        """onnx.rs:14-29 (`OnnxSession::new`)."""
        self.execution_providers: List = list(execution_providers or [])
        self.device = int(device)
        self.path = os.fspath(path)
        self._lock = threading.RLock()  # the reference serialises runs with RwLock::write (vision.rs:107)
        opts = _native.Opts(micro_batch=int(micro_batch), profile=1 if profile else 0)
        handle = C.c_void_p()
        rc = _native.lib.clipb200_engine_create(self.path.encode(), self.device, C.byref(opts), C.byref(handle))
        if rc != _native.OK:
            _raise_engine_error(rc)
        self._h = handle

    # -- introspection (onnx.rs:32-46)
    def input_names(self) -> List[str]:
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
        return self._h

    @property
    def embed_dim(self) -> int:
        return int(_native.lib.clipb200_engine_embed_dim(self._h))

    @property
    def image_size(self) -> int:
        return int(_native.lib.clipb200_engine_image_size(self._h))

    @property
    def context_length(self) -> int:
        return int(_native.lib.clipb200_engine_context_length(self._h))

    @property
    def weight_bytes(self) -> int:
        return int(_native.lib.clipb200_engine_weight_bytes(self._h))

    @property
    def launch_count(self) -> int:
        return int(_native.lib.clipb200_engine_launch_count(self._h))

    def check(self, rc: int) -> None:
        if rc != _native.OK:
            _raise_engine_error(rc)

    def synchronize(self) -> None:
        self.check(_native.lib.clipb200_engine_synchronize(self._h))

    def profile(self, reset: bool = True) -> dict:
        p = _native.Profile()
        self.check(_native.lib.clipb200_engine_profile(self._h, C.byref(p), 1 if reset else 0))
        return {"ms": {n: p.ms[i] for i, n in enumerate(_native.PROF_NAMES)},
                "launches": {n: int(p.launches[i]) for i, n in enumerate(_native.PROF_NAMES)},
                "gemm_flops": p.gemm_flops}

    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            _native.lib.clipb200_engine_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
