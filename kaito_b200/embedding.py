"""Embedding-model interface of the host (mirrors embedding/base.py:21-31 BaseEmbeddingModel).
`GpuBertEmbedding` is the production embedder for the CRD's `embedding.local.modelID` (K5, the BERT forward in
libkaito_rag); `HashingEmbedding` is a deterministic stand-in without weights for the host-logic
tests (no checkpoints are available offline) and never the default of the service."""
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


BGE_QUERY_INSTRUCTION = "Represent this question for searching relevant passages: "


class GpuBertEmbedding(BaseEmbeddingModel):
    """bge-small / bge-base / bge-large on the GPU (K5, kaito_b200/csrc/embed.cu): WordPiece on the host,
    BERT encoder forward + CLS pooling + L2 normalisation in libkaito_rag.  Mirrors
    LocalHuggingFaceEmbedding (embedding/huggingface_local_embedding.py:31-61); queries get the bge
    instruction prefix LlamaIndex adds for BAAI/bge-* models [3P-unverified]."""

    def __init__(self, engine, tokenizer, config: dict, state_dict: dict, query_instruction: str | None = BGE_QUERY_INSTRUCTION,
                 max_batch_tokens: int = 65536):
        from . import _native
        self.tokenizer, self.query_instruction, self.max_batch_tokens = tokenizer, query_instruction, max_batch_tokens
        self.dim = config["hidden_size"]
        self._emb = _native.Embedder(engine, config["num_hidden_layers"], config["hidden_size"], config["num_attention_heads"],
                                     config["intermediate_size"], config["vocab_size"], config.get("max_position_embeddings", 512),
                                     config.get("type_vocab_size", 2), config.get("layer_norm_eps", 1e-12))
        self._emb.load_state_dict(state_dict)
        import threading
        self._mu = threading.Lock()          # one forward at a time: the embedder owns one set of activation buffers

    @classmethod
    def from_pretrained(cls, engine, model_dir: str, **kw):
        """model_dir: a Hugging Face snapshot (config.json, vocab.txt, model.safetensors or pytorch_model.bin)."""
        import json
        import os
        from .text import WordPieceTokenizer
        cfg = json.load(open(os.path.join(model_dir, "config.json")))
        st_path = os.path.join(model_dir, "model.safetensors")
        if os.path.exists(st_path):
            from safetensors.numpy import load_file
            state = load_file(st_path)
        else:
            import torch
            state = {k: v.float().numpy() for k, v in torch.load(os.path.join(model_dir, "pytorch_model.bin"), map_location="cpu").items()}
        state = {k[5:] if k.startswith("bert.") else k: v for k, v in state.items()}
        return cls(engine, WordPieceTokenizer.from_file(os.path.join(model_dir, "vocab.txt")), cfg, state, **kw)

    def get_embedding_dimension(self) -> int:
        return self.dim

    def _embed_tokens(self, token_lists):
        import numpy as np
        out, batch, ntok = [], [], 0
        for t in token_lists:
            if batch and ntok + len(t) > self.max_batch_tokens:
                with self._mu:
                    out.append(self._emb.embed(batch))
                batch, ntok = [], 0
            batch.append(t); ntok += len(t)
        if batch:
            with self._mu:
                out.append(self._emb.embed(batch))
        return np.concatenate(out)

    def get_text_embedding_batch(self, texts):
        return self._embed_tokens(self.tokenizer.encode_batch(list(texts)))

    def get_text_embedding(self, text: str):
        return self.get_text_embedding_batch([text])[0]

    def get_query_embedding(self, query: str):
        return self.get_text_embedding((self.query_instruction or "") + query)

    def get_query_embedding_batch(self, queries):
        """one K5 forward for a coalesced batch of queries (kaito_b200.batcher); the token ids stay packed numpy arrays from
        the native tokeniser to the embedder's C entry point"""
        texts = [(self.query_instruction or "") + q for q in queries]
        if hasattr(self.tokenizer, "encode_batch_flat"):
            flat, offs = self.tokenizer.encode_batch_flat(texts)
            if len(flat) <= self.max_batch_tokens:
                with self._mu:
                    return self._emb.embed_flat(flat, offs)
        return self.get_text_embedding_batch(texts)


class RemoteEmbeddingModel(BaseEmbeddingModel):
    """`embedding.remote` of the CRD (embedding/remote_embedding.py:24-76): POST {"inputs": text} with a bearer token to
    REMOTE_EMBEDDING_URL, the reply is the embedding as a JSON list.  Host-side HTTP only -- the retrieval kernels take the
    vectors exactly as with the local model."""

    def __init__(self, model_url: str, api_key: str, transport=None):
        self.model_url, self.api_key, self._transport = model_url, api_key, transport
        self._client = None

    def _http(self):
        if self._client is None:
            import httpx
            kw = {"transport": self._transport} if self._transport is not None else {}
            self._client = httpx.Client(timeout=300.0, **kw)
        return self._client

    def get_text_embedding(self, text: str):
        import httpx
        headers = {"Authorization": f"Bearer {self.api_key}", "Content-Type": "application/json"}
        try:
            r = self._http().post(self.model_url, headers=headers, json={"inputs": text})
            r.raise_for_status()
            emb = r.json()
        except httpx.HTTPError as e:
            raise RuntimeError(f"Failed to get embedding from remote model: {e}")
        if not isinstance(emb, list):
            raise ValueError("Unexpected response format. Expected a list.")
        return np.asarray(emb, np.float32)

    def get_embedding_dimension(self) -> int:
        return len(self.get_text_embedding("This is a dummy sentence."))
