"""Model directory plumbing — mirrors `/root/reference/src/model_manager.rs`."""
from __future__ import annotations

import os
from pathlib import Path

from . import error

# model_manager.rs:8-18
MODEL_FILES = (
    "model_config.json",
    "open_clip_config.json",
    "special_tokens_map.json",
    "text.onnx",
    "tokenizer.json",
    "tokenizer_config.json",
    "visual.onnx",
    "text.onnx.data",
    "visual.onnx.data",
)


def get_hf_model(model_id: str) -> Path:
    """model_manager.rs:22-40 downloads the nine files from the Hub; there is no network in this build."""
    raise error.HfHub(f"Hugging Face Hub error: hf-hub support is not built in (no network); export '{model_id}' "
                      f"locally and use from_local_id / from_local_dir")


def get_default_base_folder() -> Path:
    """model_manager.rs:44-49."""
    home = os.path.expanduser("~")
    if not home or home == "~":
        return Path(".open_clip_cache")
    return Path(home) / ".cache" / "open_clip_rs"


def verify_model_dir(model_dir) -> None:
    """model_manager.rs:52-68."""
    model_dir = Path(model_dir)
    if not model_dir.exists():
        raise error.ModelFolderNotFound(model_dir)
    for file in MODEL_FILES:
        if not (model_dir / file).is_file():
            raise error.MissingModelFile(model_dir, file)
