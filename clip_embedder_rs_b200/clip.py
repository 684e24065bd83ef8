"""`Clip` — mirrors `/root/reference/src/clip.rs:14-186`: vision + text embedders and the similarity tail
(dot product, `mul_add(logit_scale, logit_bias)`, softmax / sigmoid on the GPU; the stable descending sort stays on
the host as in clip.rs:129,167)."""
from __future__ import annotations

import functools
import math
from pathlib import Path
from typing import List, Sequence, Tuple

import numpy as np

from . import _native, error, model_manager
from .config import ModelConfig
from .text import TextEmbedder
from .vision import VisionEmbedder, _Builder, _HfBuilder, _IdBuilder


def _cmp_desc(a, b) -> int:
    """`b.1.partial_cmp(&a.1).unwrap_or(Ordering::Equal)` (clip.rs:129)."""
    x, y = a[1], b[1]
    if math.isnan(x) or math.isnan(y):
        return 0
    return -1 if y < x else (1 if y > x else 0)


class Clip:
    vision: VisionEmbedder
    text: TextEmbedder
    model_dir: Path

    @classmethod
    def from_hf(cls, model_id: str) -> _Builder:  # clip.rs:23-34
        return _HfBuilder(cls, model_id)

    @classmethod
    def from_local_id(cls, model_id: str) -> _IdBuilder:  # clip.rs:37-48
        return _IdBuilder(cls, model_id)

    @classmethod
    def from_local_dir(cls, model_dir) -> _Builder:  # clip.rs:51-66
        return _Builder(cls, Path(model_dir))

    @classmethod
    def _load(cls, model_dir: Path, execution_providers=None, **kw) -> "Clip":
        model_manager.verify_model_dir(model_dir)
        self = cls.__new__(cls)
        self.vision = VisionEmbedder._load(Path(model_dir), execution_providers, **kw)
        self.text = TextEmbedder._load(Path(model_dir), execution_providers, **kw)
        self.model_dir = Path(model_dir)
        self._kw = kw
        return self

    def duplicate(self) -> "Clip":  # clip.rs:69-73
        return type(self)._load(self.model_dir, self.vision.session.execution_providers, **self._kw)

    def get_model_config(self) -> ModelConfig:  # clip.rs:75-77
        return self.text.model_config

    # ------------------------------------------------------------------------------------------ similarity tail
    def _probabilities(self, embs: np.ndarray, query: np.ndarray, raw_logits: bool = False) -> np.ndarray:
        mc = self.text.model_config
        scale = 1.0 if mc.logit_scale is None else float(mc.logit_scale)
        bias = 0.0 if mc.logit_bias is None else float(mc.logit_bias)
        activation = mc.activation_function or "softmax"
        act = _native.ACT_SIGMOID if activation == "sigmoid" else _native.ACT_SOFTMAX
        embs = np.ascontiguousarray(embs, dtype=np.float32)
        query = np.ascontiguousarray(query, dtype=np.float32)
        probs = np.empty(embs.shape[0], dtype=np.float32)
        rc = _native.lib.clipb200_similarity(self.text.session.device, embs.ctypes.data, query.ctypes.data,
                                             embs.shape[0], embs.shape[1], scale, bias, act, probs.ctypes.data)
        if rc != _native.OK:
            raise error.Ort(_native.last_error(), rc)
        return probs

    def compare(self, image, text: str) -> float:  # clip.rs:81-90
        vision_emb = self.vision.embed_image(image)
        text_emb = self.text.embed_text(text)
        mc = self.text.model_config
        scale = 1.0 if mc.logit_scale is None else float(mc.logit_scale)
        bias = 0.0 if mc.logit_bias is None else float(mc.logit_bias)
        # a single logit: sigmoid^-1 is not needed, ask the tail for the sigmoid input by using scale/bias directly
        probs = np.empty(1, dtype=np.float32)
        v = np.ascontiguousarray(vision_emb.reshape(1, -1), dtype=np.float32)
        t = np.ascontiguousarray(text_emb, dtype=np.float32)
        rc = _native.lib.clipb200_similarity(self.text.session.device, v.ctypes.data, t.ctypes.data, 1, v.shape[1],
                                             scale, bias, 2, probs.ctypes.data)
        if rc != _native.OK:
            raise error.Ort(_native.last_error(), rc)
        return float(probs[0])

    def classify(self, image, labels: Sequence[str]) -> List[Tuple[str, float]]:  # clip.rs:94-132
        vision_emb = self.vision.embed_image(image)
        text_embs = self.text.embed_texts(labels)
        probs = self._probabilities(text_embs, vision_emb)
        results = [(str(l), float(p)) for l, p in zip(labels, probs)]
        return sorted(results, key=functools.cmp_to_key(_cmp_desc))

    def rank_images(self, images: Sequence, text: str) -> List[Tuple[int, float]]:  # clip.rs:136-170
        img_embs = self.vision.embed_images(images)
        text_emb = self.text.embed_text(text)
        probs = self._probabilities(img_embs, text_emb)
        results = [(i, float(p)) for i, p in enumerate(probs)]
        return sorted(results, key=functools.cmp_to_key(_cmp_desc))

    @staticmethod
    def softmax(logits: Sequence[float]) -> List[float]:  # clip.rs:174-179
        lg = np.ascontiguousarray(logits, dtype=np.float32).reshape(-1, 1)
        one = np.ones(1, dtype=np.float32)
        probs = np.empty(lg.shape[0], dtype=np.float32)
        rc = _native.lib.clipb200_similarity(0, lg.ctypes.data, one.ctypes.data, lg.shape[0], 1, 1.0, 0.0,
                                             _native.ACT_SOFTMAX, probs.ctypes.data)
        if rc != _native.OK:
            raise error.Ort(_native.last_error(), rc)
        return [float(p) for p in probs]

    @staticmethod
    def sigmoid(logit: float) -> float:  # clip.rs:183-185
        lg = np.asarray([[logit]], dtype=np.float32)
        one = np.ones(1, dtype=np.float32)
        probs = np.empty(1, dtype=np.float32)
        rc = _native.lib.clipb200_similarity(0, lg.ctypes.data, one.ctypes.data, 1, 1, 1.0, 0.0,
                                             _native.ACT_SIGMOID, probs.ctypes.data)
        if rc != _native.OK:
            raise error.Ort(_native.last_error(), rc)
        return float(probs[0])
