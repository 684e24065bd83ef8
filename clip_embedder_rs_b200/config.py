"""`ModelConfig` / `OpenClipConfig` — mirrors `/root/reference/src/config.rs:7-64` (same fields, same defaults,
unknown JSON keys ignored like serde does)."""
from __future__ import annotations

import json
from dataclasses import dataclass
from typing import List, Optional

from . import error


def _read_json(path):
    try:
        with open(path, "r", encoding="utf-8") as f:
            text = f.read()
    except OSError as e:
        raise error.Io(f"IO error: {e}") from e
    try:
        return json.loads(text)
    except json.JSONDecodeError as e:
        raise error.Json(f"JSON error: {e}") from e


def _req(d: dict, key: str, where: str):
    if not isinstance(d, dict) or key not in d:
        raise error.Json(f"JSON error: missing field `{key}` in {where}")
    return d[key]


@dataclass
class ModelConfig:  # config.rs:7-14
    tokenizer_needs_lowercase: bool = False
    activation_function: Optional[str] = None
    logit_scale: Optional[float] = None
    logit_bias: Optional[float] = None
    pad_id: Optional[int] = None

    @classmethod
    def from_file(cls, path) -> "ModelConfig":
        d = _read_json(path)
        return cls(bool(d.get("tokenizer_needs_lowercase", False)), d.get("activation_function"),
                   d.get("logit_scale"), d.get("logit_bias"), d.get("pad_id"))


@dataclass
class VisionCfg:  # config.rs:37-42
    image_size: int
    layers: Optional[int] = None
    width: Optional[int] = None


@dataclass
class TextCfg:  # config.rs:44-48
    context_length: int
    hf_tokenizer_name: Optional[str] = None


@dataclass
class ModelCfg:  # config.rs:30-35
    embed_dim: int
    vision_cfg: VisionCfg
    text_cfg: TextCfg


@dataclass
class PreprocessCfg:  # config.rs:50-57; defaults :59-64
    mean: List[float]
    std: List[float]
    interpolation: str = "bicubic"
    resize_mode: str = "shortest"


@dataclass
class OpenClipConfig:  # config.rs:24-28
    model_cfg: ModelCfg
    preprocess_cfg: PreprocessCfg

    @classmethod
    def from_file(cls, path) -> "OpenClipConfig":
        d = _read_json(path)
        mc = _req(d, "model_cfg", "open_clip_config.json")
        pc = _req(d, "preprocess_cfg", "open_clip_config.json")
        vc = _req(mc, "vision_cfg", "model_cfg")
        tc = _req(mc, "text_cfg", "model_cfg")
        layers = vc.get("layers")
        mean, std = _req(pc, "mean", "preprocess_cfg"), _req(pc, "std", "preprocess_cfg")
        if len(mean) != 3 or len(std) != 3:
            raise error.Json("JSON error: preprocess_cfg.mean/std must have 3 entries")
        return cls(
            ModelCfg(int(_req(mc, "embed_dim", "model_cfg")),
                     VisionCfg(int(_req(vc, "image_size", "vision_cfg")),
                               layers if isinstance(layers, int) else None, vc.get("width")),
                     TextCfg(int(_req(tc, "context_length", "text_cfg")), tc.get("hf_tokenizer_name"))),
            PreprocessCfg([float(x) for x in mean], [float(x) for x in std],
                          pc.get("interpolation", "bicubic"), pc.get("resize_mode", "shortest")))
