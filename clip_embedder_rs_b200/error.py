"""`ClipError` — mirrors the reference's error enum (`/root/reference/src/error.rs:9-41`).

Each Rust variant is a subclass so callers can match on the kind; `Ort` carries the engine's last-error string
the same way `From<ort::Error>` stringifies ORT errors (`src/error.rs:62-66`)."""
from __future__ import annotations


class ClipError(Exception):
    pass


class Io(ClipError):
    pass


class Json(ClipError):
    pass


class Ort(ClipError):
    """Engine failure (the reference's `ClipError::Ort(String)`); `.code` is the C-ABI status."""

    def __init__(self, msg: str, code: int = 0):
        super().__init__(f"ONNX Runtime Error: {msg}")
        self.code = code


class Image(ClipError):
    pass


class Tokenizer(ClipError):
    pass


class Config(ClipError):
    def __init__(self, msg: str):
        super().__init__(f"Configuration error: {msg}")


class Inference(ClipError):
    def __init__(self, msg: str):
        super().__init__(f"Inference error: {msg}")


class Shape(ClipError):
    pass


class ModelFolderNotFound(ClipError):
    def __init__(self, path):
        super().__init__(f"Model folder not found, generate it with `uv run pull_onnx.py -h`. '{path}'")
        self.path = path


class HfHub(ClipError):
    pass


class MissingModelFile(ClipError):
    def __init__(self, model_dir, file):
        super().__init__(f"Missing model file '{file}' in folder '{model_dir}'")
        self.model_dir, self.file = model_dir, file


class Resize(ClipError):
    pass
