"""clip_embedder_rs_b200 — host-side mirror of `open_clip_inference` (RuurdBijlsma/clip-embedder-rs) over the
B200-native engine `libclipb200.so`.  Same names as the Rust crate's re-exports (`/root/reference/src/lib.rs:170-181`).
"""
from . import error
from .clip import Clip
from .config import ModelConfig, OpenClipConfig
from .error import ClipError
from .onnx import OnnxSession
from .text import TextEmbedder
from .vision import VisionEmbedder

__all__ = ["Clip", "VisionEmbedder", "TextEmbedder", "OnnxSession", "ClipError", "ModelConfig", "OpenClipConfig",
           "error"]
