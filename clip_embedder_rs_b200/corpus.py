"""Corpus search: `Clip::rank_images` (`/root/reference/src/clip.rs:136-170`) for corpora that do not fit one
`embed_images` call (BASELINE config 5: 100k images).  Embeddings are appended to an HBM-resident matrix as they
are produced; a query is one fused similarity pass over the matrix (dot -> mul_add(scale, bias) -> softmax over the
whole corpus, or sigmoid) and the stable descending sort of clip.rs:167 on the host; many queries at once go through
one tcgen05 GEMM pass over the corpus and a per-query top-k on the GPU (`search_corpus`)."""
from __future__ import annotations

import ctypes as C
import functools
from typing import List, Sequence, Tuple

import numpy as np

from . import _native, error
from .clip import Clip, _cmp_desc


class EmbeddingCorpus:
    def __init__(self, dim: int, capacity: int, device: int = 0):
        h = C.c_void_p()
        rc = _native.lib.clipb200_corpus_create(int(device), int(dim), int(capacity), C.byref(h))
        if rc != _native.OK:
            raise error.Ort(_native.last_error(), rc)
        self._h, self.dim, self.capacity = h, int(dim), int(capacity)

    def __len__(self) -> int:
        return int(_native.lib.clipb200_corpus_size(self._h))

    def append(self, rows: np.ndarray) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise error.Shape(f"Shape error: expected [n,{self.dim}], got {rows.shape}")
        rc = _native.lib.clipb200_corpus_append(self._h, rows.ctypes.data, rows.shape[0])
        if rc != _native.OK:
            raise error.Ort(_native.last_error(), rc)

    def probabilities(self, query: np.ndarray, scale: float, bias: float, sigmoid: bool) -> np.ndarray:
        q = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
        if q.shape[0] != self.dim:
            raise error.Shape(f"Shape error: query has {q.shape[0]} dims, corpus {self.dim}")
        probs = np.empty(len(self), dtype=np.float32)
        rc = _native.lib.clipb200_corpus_rank(self._h, q.ctypes.data, float(scale), float(bias),
                                              _native.ACT_SIGMOID if sigmoid else _native.ACT_SOFTMAX, probs.ctypes.data)
        if rc != _native.OK:
            raise error.Ort(_native.last_error(), rc)
        return probs

    def search(self, queries: np.ndarray, k: int, scale: float, bias: float, sigmoid: bool) -> Tuple[np.ndarray, np.ndarray]:
        """Top-k of every query over the resident corpus, ranked on the GPU: (index int64 [Q,k], prob f32 [Q,k])."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise error.Shape(f"Shape error: queries must be [Q,{self.dim}], got {q.shape}")
        idx = np.empty((q.shape[0], int(k)), dtype=np.int64)
        prob = np.empty((q.shape[0], int(k)), dtype=np.float32)
        rc = _native.lib.clipb200_corpus_search(self._h, q.ctypes.data, q.shape[0], int(k), float(scale), float(bias),
                                                _native.ACT_SIGMOID if sigmoid else _native.ACT_SOFTMAX,
                                                idx.ctypes.data, prob.ctypes.data)
        if rc != _native.OK:
            raise error.Ort(_native.last_error(), rc)
        return idx, prob

    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            _native.lib.clipb200_corpus_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rank_corpus(clip: Clip, corpus: EmbeddingCorpus, text: str, top_k: int = 0) -> List[Tuple[int, float]]:
    """clip.rs:136-170 against a resident corpus; `top_k` = 0 returns the full ranking."""
    mc = clip.text.model_config
    probs = corpus.probabilities(clip.text.embed_text(text), 1.0 if mc.logit_scale is None else mc.logit_scale,
                                 0.0 if mc.logit_bias is None else mc.logit_bias,
                                 (mc.activation_function or "softmax") == "sigmoid")
    results = sorted(((i, float(p)) for i, p in enumerate(probs)), key=functools.cmp_to_key(_cmp_desc))
    return results[:top_k] if top_k else results


def search_corpus(clip: Clip, corpus: EmbeddingCorpus, texts: Sequence[str], top_k: int) -> List[List[Tuple[int, float]]]:
    """`rank_images` (clip.rs:136-170) for many texts at once: one GEMM pass over the corpus and a GPU top-k; entry q is
    what `rank_corpus(clip, corpus, texts[q], top_k)` returns."""
    mc = clip.text.model_config
    idx, prob = corpus.search(clip.text.embed_texts(texts), top_k, 1.0 if mc.logit_scale is None else mc.logit_scale,
                              0.0 if mc.logit_bias is None else mc.logit_bias,
                              (mc.activation_function or "softmax") == "sigmoid")
    return [[(int(i), float(p)) for i, p in zip(idx[q], prob[q])] for q in range(len(texts))]
