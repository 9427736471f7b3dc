"""Embedding-model interface of the host (mirrors embedding/base.py:21-31 BaseEmbeddingModel) and a
deterministic hashing embedder used by tests and the demo service until the GPU BERT forward (K5) lands.
The production embedder for the CRD's `embedding.local.modelID` is K5."""
from __future__ import annotations

import hashlib
import re

import numpy as np


class BaseEmbeddingModel:
    def get_text_embedding(self, text: str) -> list[float]:
        raise NotImplementedError

    def get_text_embedding_batch(self, texts: list[str]) -> np.ndarray:
        return np.asarray([self.get_text_embedding(t) for t in texts], np.float32)

    def get_query_embedding(self, query: str) -> list[float]:
        return self.get_text_embedding(query)

    def get_embedding_dimension(self) -> int:
        """reference probes by embedding a dummy sentence (huggingface_local_embedding.py:55-61)"""
        return len(self.get_text_embedding("dimension probe"))


class HashingEmbedding(BaseEmbeddingModel):
    """Unit-norm sum of per-word pseudo-random Gaussian vectors (seeded by sha256 of the word):
    deterministic, no weights, texts sharing words are close.  NOT a language model."""

    def __init__(self, dim: int = 384):
        self.dim = dim
        self._cache: dict[str, np.ndarray] = {}

    def _word(self, w: str) -> np.ndarray:
        v = self._cache.get(w)
        if v is None:
            seed = int.from_bytes(hashlib.sha256(w.encode()).digest()[:8], "little")
            v = np.random.default_rng(seed).standard_normal(self.dim).astype(np.float32)
            self._cache[w] = v
        return v

    def get_text_embedding(self, text: str):
        words = re.findall(r"\w+", text.lower()) or [""]
        v = np.sum([self._word(w) for w in words], axis=0)
        n = float(np.linalg.norm(v))
        return (v / n if n > 0 else v).astype(np.float32)

    def get_embedding_dimension(self) -> int:
        return self.dim
