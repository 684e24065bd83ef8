"""`TextEmbedder` — mirrors `/root/reference/src/text.rs:14-170`.

Tokenisation stays on the host and uses the same Rust crate as the reference (`tokenizers` 0.22.2, through its
Python binding), configured exactly as text.rs:70-85 does, so token ids are bit-exact by construction."""
from __future__ import annotations

from pathlib import Path
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _native, error, model_manager
from .config import ModelConfig, OpenClipConfig
from .onnx import OnnxSession
from .vision import _Builder, _HfBuilder, _IdBuilder


class TextEmbedder:
    session: OnnxSession
    config: OpenClipConfig
    model_config: ModelConfig
    model_dir: Path

    @classmethod
    def from_hf(cls, model_id: str) -> _Builder:  # text.rs:27-38
        return _HfBuilder(cls, model_id)

    @classmethod
    def from_local_id(cls, model_id: str) -> _IdBuilder:  # text.rs:41-52
        return _IdBuilder(cls, model_id)

    @classmethod
    def from_local_dir(cls, model_dir) -> _Builder:  # text.rs:55-101
        return _Builder(cls, Path(model_dir))

    @classmethod
    def _load(cls, model_dir: Path, execution_providers=None, device: int = 0, micro_batch: int = 0,
              profile: bool = False, devices=None) -> "TextEmbedder":
        try:
            from tokenizers import Tokenizer
        except ImportError as e:  # pragma: no cover
            raise error.Tokenizer(f"Tokenization error: {e}") from e
        model_manager.verify_model_dir(model_dir)
        self = cls.__new__(cls)
        self.model_config = ModelConfig.from_file(model_dir / "model_config.json")
        self.session = OnnxSession(model_dir / "text.onnx", execution_providers, device, micro_batch, profile, devices)
        self.config = OpenClipConfig.from_file(model_dir / "open_clip_config.json")
        try:
            tokenizer = Tokenizer.from_file(str(model_dir / "tokenizer.json"))
        except Exception as e:
            raise error.Tokenizer(f"Tokenization error: {e}") from e
        pad_id = self.model_config.pad_id
        if pad_id is None:
            pad_id = tokenizer.get_vocab(True).get("<pad>")  # text.rs:70-73
        if pad_id is None:
            raise error.Config("No pad token found in tokenizer")
        ctx_len = self.config.model_cfg.text_cfg.context_length
        tokenizer.enable_padding(length=ctx_len, pad_id=int(pad_id))  # PaddingStrategy::Fixed, text.rs:76-81
        tokenizer.enable_truncation(max_length=ctx_len)  # text.rs:82-85
        self._tokenizer = tokenizer
        id_name = self.session.find_input(["input_ids"])
        if id_name is None:
            raise error.Config("Could not find text input node")
        self._id_name = id_name
        self._mask_name: Optional[str] = self.session.find_input(["attention_mask"])
        self.model_dir = Path(model_dir)
        self._kw = dict(device=device, micro_batch=micro_batch, profile=profile, devices=devices)
        return self

    def duplicate(self) -> "TextEmbedder":  # text.rs:104-108
        return type(self)._load(self.model_dir, self.session.execution_providers, **self._kw)

    def tokenize(self, texts: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:  # text.rs:111-139
        texts = [str(t) for t in texts]
        if self.model_config.tokenizer_needs_lowercase:
            texts = [t.lower() for t in texts]
        try:
            enc = self._tokenizer.encode_batch(texts, add_special_tokens=True)
        except Exception as e:
            raise error.Tokenizer(f"Tokenization error: {e}") from e
        seq_len = self.config.model_cfg.text_cfg.context_length
        ids = np.fromiter((i for e in enc for i in e.ids), dtype=np.int64)
        mask = np.fromiter((i for e in enc for i in e.attention_mask), dtype=np.int64)
        if ids.size != len(enc) * seq_len:
            raise error.Shape(f"Shape error: tokenizer produced {ids.size} ids for {len(enc)}x{seq_len}")
        return ids.reshape(len(enc), seq_len), mask.reshape(len(enc), seq_len)

    def embed_text(self, text: str) -> np.ndarray:  # text.rs:142-146
        return self.embed_texts([text]).reshape(-1)

    def embed_texts(self, texts: Sequence[str]) -> np.ndarray:  # text.rs:150-169
        ids, mask = self.tokenize(texts)
        return self.embed_ids(ids, mask if self._mask_name else None)

    def embed_ids(self, ids: np.ndarray, mask: Optional[np.ndarray] = None) -> np.ndarray:
        """`session.run(inputs![input_ids (, attention_mask)])` (text.rs:153-166)."""
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        if ids.ndim != 2:
            raise error.Shape(f"Shape error: input_ids must be [B,ctx], got {ids.shape}")
        if ids.shape[0] == 0:
            raise error.Inference("Empty batch")
        out = np.empty((ids.shape[0], self.session.embed_dim), dtype=np.float32)
        mptr = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.int64)
            mptr = mask.ctypes.data
        self.session.run_ids(ids.ctypes.data, mptr, ids.shape[0], ids.shape[1], out.ctypes.data)
        return out
