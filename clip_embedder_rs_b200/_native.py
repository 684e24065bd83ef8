"""ctypes binding of `libclipb200.so` (C ABI declared in `include/clipb200.h`).

This is the Python twin of the `extern "C"` block a Rust maintainer would add to the reference (see
INTEGRATION.md).  There is no fallback: if the shared library is missing or cannot be loaded, importing this
module raises, and every engine call fails loudly when no sm_100 GPU is present.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

OK, ERR_INVALID_ARG, ERR_IO, ERR_PARSE, ERR_CUDA, ERR_UNSUPPORTED = 0, 1, 2, 3, 4, 5
KIND_VISION, KIND_TEXT = 0, 1
ACT_SOFTMAX, ACT_SIGMOID = 0, 1
PROF_CLASSES = 8
PROF_NAMES = ("gemm", "attention", "layernorm", "preprocess", "misc", "h2d", "d2h", "dwconv")


class Opts(C.Structure):
    _fields_ = [("micro_batch", C.c_int32), ("profile", C.c_int32), ("reserved", C.c_int32 * 6)]


class Preproc(C.Structure):
    _fields_ = [("mean", C.c_float * 3), ("std", C.c_float * 3), ("interpolation", C.c_int32),
                ("resize_mode", C.c_int32)]


class Profile(C.Structure):
    _fields_ = [("ms", C.c_double * PROF_CLASSES), ("launches", C.c_int64 * PROF_CLASSES),
                ("gemm_flops", C.c_double), ("conv_bytes", C.c_double)]


LIB_PATH = Path(__file__).resolve().parent / "libclipb200.so"

# name -> (restype, argtypes); must list every symbol include/clipb200.h declares (tests check this).
SIGNATURES = {
    "clipb200_engine_create": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(Opts), C.POINTER(C.c_void_p)]),
    "clipb200_engine_destroy": (None, [C.c_void_p]),
    "clipb200_last_error": (C.c_char_p, []),
    "clipb200_version": (C.c_char_p, []),
    "clipb200_onnx_inspect": (C.c_int, [C.c_char_p, C.c_char_p, C.c_size_t]),
    "clipb200_onnx_read_tensor": (C.c_int, [C.c_char_p, C.c_char_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "clipb200_engine_num_inputs": (C.c_int, [C.c_void_p]),
    "clipb200_engine_input_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "clipb200_engine_kind": (C.c_int, [C.c_void_p]),
    "clipb200_engine_embed_dim": (C.c_int64, [C.c_void_p]),
    "clipb200_engine_image_size": (C.c_int64, [C.c_void_p]),
    "clipb200_engine_context_length": (C.c_int64, [C.c_void_p]),
    "clipb200_engine_weight_bytes": (C.c_int64, [C.c_void_p]),
    "clipb200_vision_embed_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "clipb200_vision_embed_rgb8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                             C.POINTER(Preproc), C.c_void_p]),
    "clipb200_vision_embed_rgb8_var": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                                 C.POINTER(Preproc), C.c_void_p]),
    "clipb200_resize_rgb8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(Preproc), C.c_void_p]),
    "clipb200_preprocess_rgb8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                           C.POINTER(Preproc), C.c_void_p]),
    "clipb200_text_embed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "clipb200_similarity": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_float,
                                      C.c_int, C.c_void_p]),
    "clipb200_corpus_create": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]),
    "clipb200_corpus_destroy": (None, [C.c_void_p]),
    "clipb200_corpus_append": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "clipb200_corpus_size": (C.c_int64, [C.c_void_p]),
    "clipb200_corpus_rank": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_void_p]),
    "clipb200_corpus_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_int,
                                         C.c_void_p, C.c_void_p]),
    "clipb200_pool_create": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int32, C.POINTER(Opts), C.POINTER(C.c_void_p)]),
    "clipb200_pool_destroy": (None, [C.c_void_p]),
    "clipb200_pool_size": (C.c_int, [C.c_void_p]),
    "clipb200_pool_device": (C.c_int, [C.c_void_p, C.c_int]),
    "clipb200_pool_kind": (C.c_int, [C.c_void_p]),
    "clipb200_pool_embed_dim": (C.c_int64, [C.c_void_p]),
    "clipb200_pool_image_size": (C.c_int64, [C.c_void_p]),
    "clipb200_pool_context_length": (C.c_int64, [C.c_void_p]),
    "clipb200_pool_num_inputs": (C.c_int, [C.c_void_p]),
    "clipb200_pool_input_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "clipb200_pool_launch_count": (C.c_int64, [C.c_void_p]),
    "clipb200_pool_vision_embed_rgb8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                                  C.POINTER(Preproc), C.c_void_p]),
    "clipb200_pool_vision_embed_rgb8_var": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                                      C.POINTER(Preproc), C.c_void_p]),
    "clipb200_pool_text_embed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "clipb200_vision_embed_rgb8_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(Preproc),
                                                    C.c_void_p]),
    "clipb200_text_embed_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "clipb200_host_alloc": (C.c_void_p, [C.c_size_t]),
    "clipb200_host_free": (None, [C.c_void_p]),
    "clipb200_device_alloc": (C.c_void_p, [C.c_int, C.c_size_t]),
    "clipb200_device_free": (None, [C.c_int, C.c_void_p]),
    "clipb200_memcpy_h2d": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "clipb200_memcpy_d2h": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "clipb200_device_count": (C.c_int, []),
    "clipb200_engine_record_event": (C.c_int, [C.c_void_p, C.c_int]),
    "clipb200_engine_elapsed_ms": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "clipb200_engine_synchronize": (C.c_int, [C.c_void_p]),
    "clipb200_engine_profile": (C.c_int, [C.c_void_p, C.POINTER(Profile), C.c_int]),
    "clipb200_engine_launch_count": (C.c_int64, [C.c_void_p]),
    "clipb200_engine_flush_l2": (C.c_int, [C.c_void_p]),
}


def _load() -> C.CDLL:
    if not LIB_PATH.is_file():
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA engine first (python -c 'import __graft_entry__ as g; g.build()' "
            f"or make -C clip_embedder_rs_b200/csrc). There is no CPU fallback.")
    lib = C.CDLL(os.fspath(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def last_error() -> str:
    msg = lib.clipb200_last_error()
    return msg.decode("utf-8", "replace") if msg else ""
