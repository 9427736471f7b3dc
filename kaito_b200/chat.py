"""`/v1/chat/completions`: dense kNN on the GPU (K1/K2) -> context selection -> LLM proxy.

Host logic of SURVEY.md section 8 (f3), restating
  * `BaseVectorStore.chat_completion` (presets/ragengine/vector_store/base.py:180-477): validation, pass-through rules,
    split of the messages into chat history + user prompt, `top_k = max(100, (window - prompt) / 500)`;
  * `ContextSelectionProcessor._postprocess_nodes` (vector_store/node_processors/contex_selection_node_processor.py:67-121):
    distance-ascending order, `score > 0.85` dropped, greedy token-budget packing;
  * `Inference` (inference/inference.py:166-317, 422-521): request shape sent to the LLM, pass-through, default model from
    `/v1/models`, token counting (tiktoken, `len/3` when the BPE table cannot be loaded), error envelopes.
The chat engine between them is LlamaIndex's ContextChatEngine [3P, un-vendored]: one system message built from
`DEFAULT_CONTEXT_TEMPLATE` with the selected nodes' `MetadataMode.LLM` text, then the chat history, then the user prompt;
with no node left it answers "Empty Response" without calling the LLM, which the reference turns into a pass-through
(base.py:404-412) -- restated here as such [3P-unverified wording of the template].
"""
from __future__ import annotations

import json
import os
import time
import uuid
from urllib.parse import urljoin, urlparse

import numpy as np

from .vector_store import HTTPException, embed_text

ADDITION_PROMPT_TOKENS = 150            # contex_selection_node_processor.py:22-24
DEFAULT_HTTP_TIMEOUT = 300.0            # inference.py:56
CONTEXT_TEMPLATE = ("Use the context information below to assist the user."
                    "\n--------------------\n{context_str}\n--------------------\n")
MAX_DENSE_K = 1024                      # KRAG_MAX_POOL (include/kaito_rag.h)


def chat_config() -> dict:
    """config.py:64-69, 115-127"""
    g = os.getenv
    return {
        "llm_inference_url": g("LLM_INFERENCE_URL"),
        "llm_access_secret": g("LLM_ACCESS_SECRET", "default-access-secret"),
        "llm_context_window": int(g("LLM_CONTEXT_WINDOW", 64000)),
        "similarity_threshold": float(g("RAG_SIMILARITY_THRESHOLD", 0.85)),
        "context_token_fill_ratio": float(g("RAG_CONTEXT_TOKEN_FILL_RATIO", 0.5)),
        "node_token_approximation": float(g("RAG_DOCUMENT_NODE_TOKEN_APPROXIMATION", 500)),
    }


class LLMClient:
    """The `Inference` LLM of the reference: an OpenAI-compatible endpoint behind LLM_INFERENCE_URL."""

    def __init__(self, url: str | None, access_secret: str = "default-access-secret", context_window: int = 64000,
                 transport=None):
        self.url, self.context_window = url, context_window
        self.headers = {"Authorization": f"Bearer {access_secret}", "Content-Type": "application/json",
                        "User-Agent": "KAITO-RagEngine/b200"}
        self._transport = transport           # httpx transport override (tests)
        self._client = None
        self._default_model = None
        self._default_max_model_len = None
        self._model_retrieval_attempted = False
        self._token_encoder = None
        self._encoder_failed = False
        self.last_usage = None

    # ---- metadata (inference.py:470-490)
    @property
    def is_chat_model(self) -> bool:
        return bool(self.url) and "/chat/completions" in urlparse(self.url).path.lower()

    def _http(self):
        if self._client is None:
            import httpx
            kw = {"transport": self._transport} if self._transport is not None else {}
            self._client = httpx.Client(timeout=DEFAULT_HTTP_TIMEOUT, headers=self.headers, **kw)
        return self._client

    def close(self):
        if self._client is not None:
            self._client.close()
            self._client = None

    # ---- default model (inference.py:385-421)
    def _get_default_model_info(self):
        if not self._default_model and not self._model_retrieval_attempted:
            self._model_retrieval_attempted = True
            try:
                p = urlparse(self.url)
                r = self._http().get(urljoin(f"{p.scheme}://{p.netloc}", "/v1/models"))
                r.raise_for_status()
                models = r.json().get("data", [])
                if models:
                    self._default_model, self._default_max_model_len = models[0].get("id"), models[0].get("max_model_len")
            except Exception:
                pass                           # '"model" parameter will not be included with inference call'
        return self._default_model, self._default_max_model_len

    # ---- token counting (inference.py:497-521)
    def count_tokens(self, prompt: str) -> int:
        if self._token_encoder is None and not self._encoder_failed:
            try:
                import tiktoken
                model, _ = self._get_default_model_info() if self.url else (None, None)
                self._token_encoder = (tiktoken.encoding_for_model(model) if model and "gpt" in model
                                       else tiktoken.get_encoding("o200k_base"))
            except Exception:
                self._encoder_failed = True    # no BPE table (offline image): character-count fallback below
        if self._token_encoder is not None:
            try:
                return len(self._token_encoder.encode(prompt))
            except Exception:
                pass
        return int(len(prompt) / 3)

    # ---- POST helpers
    def _post_raw(self, data: dict) -> dict:
        if not self.url:
            raise HTTPException(503, "LLM inference service is not configured. Please set LLM_INFERENCE_URL environment variable.")
        r = self._http().post(self.url, json=data, headers=self.headers)
        r.raise_for_status()
        return r.json()

    def chat(self, messages: list[dict], params: dict | None = None, max_tokens: int | None = None) -> dict:
        """inference.py:166-268 (`achat`): messages = [{"role", "content"}]; returns the raw LLM JSON."""
        import httpx
        params = dict(params or {})
        try:
            base_model, _ = self._get_default_model_info()
            approx = sum(self.count_tokens(m["content"]) for m in messages if m.get("content"))
            if approx > self.context_window:
                raise HTTPException(400, f"Content length exceeds context window size ({self.context_window}). "
                                         "Please reduce the length of the messages.")
            if max_tokens is not None:
                if max_tokens > self.context_window:
                    raise HTTPException(400, f"Provided max_tokens ({max_tokens}) exceeds context window size "
                                             f"({self.context_window}). Adjusting to fit within context window.")
                max_tokens = min(max_tokens, self.context_window - approx)
            if max_tokens and max_tokens < 0:
                raise HTTPException(400, f"Provided content length exceeds max_tokens limit ({max_tokens}). Please reduce "
                                         "the length of the messages or increase max_tokens.")
            req = {"model": params.get("model", base_model), "max_tokens": max_tokens,
                   "messages": [{"role": m["role"], "content": m["content"] if isinstance(m["content"], str) else json.dumps(m["content"])}
                                for m in messages if m.get("content") is not None and m.get("content") != ""]}
            for k, v in params.items():
                if k not in req:
                    req[k] = v
            resp = self._post_raw(req)
            self.last_usage = resp.get("usage")
            return resp
        except HTTPException:
            raise
        except httpx.HTTPStatusError as e:
            raise HTTPException(500, f"An unexpected error occurred: {e}")
        except Exception as e:
            raise HTTPException(500, f"An unexpected error occurred: {e}")

    def chat_completions_passthrough(self, request: dict) -> dict:
        """inference.py:270-317: forward the caller's request untouched; `source_nodes: null` marks the pass-through."""
        import httpx
        try:
            if not self.url or "/chat/completions" not in self.url:
                raise HTTPException(400, f"Chat completions not supported through endpoint {self.url}.")
            r = self._http().post(self.url, json=request, headers=self.headers)
            r.raise_for_status()
            out = r.json()
            out["source_nodes"] = None
            return out
        except HTTPException:
            raise
        except httpx.HTTPStatusError as e:
            raise HTTPException(e.response.status_code, f"{e.response.content!s}")
        except httpx.RequestError as e:
            raise HTTPException(500, f"Error during POST request: {e}")
        except Exception as e:
            raise HTTPException(500, f"Error during POST request: {e}")


def select_context(nodes: list[tuple[object, float]], query_str: str, llm: LLMClient, fill_ratio: float,
                   max_tokens: int | None, similarity_threshold: float | None) -> list[tuple[object, float]]:
    """contex_selection_node_processor.py:67-121.  nodes: (node, L2^2 distance); node.text is what is counted."""
    if not nodes:
        return []
    budget = llm.context_window - llm.count_tokens(query_str) - ADDITION_PROMPT_TOKENS
    budget = min(max_tokens or llm.context_window, budget)
    budget = int(budget * fill_ratio)
    if budget <= 0:
        return []
    out = []
    for node, score in sorted(nodes, key=lambda x: x[1] or 0.0):          # faiss scores are distances: nearest first
        if similarity_threshold is not None and score > similarity_threshold:
            continue
        n_tok = llm.count_tokens(node.text)
        if n_tok > budget:
            continue                                                       # a later, shorter node may still fit
        budget -= n_tok
        out.append((node, score))
    return out


def _message_text(content) -> str:
    """messages_to_llamaindex (base.py:305): text of a message; list content = its text parts joined"""
    if content is None:
        return ""
    if isinstance(content, str):
        return content
    parts = []
    for p in content:
        parts.append(p if isinstance(p, str) else p.get("text", ""))
    return "\n".join(parts)


def chat_completion(store, llm: LLMClient, request: dict, cfg: dict | None = None) -> dict:
    """BaseVectorStore.chat_completion (base.py:180-477) over the CUDA engine."""
    cfg = cfg or chat_config()
    index_name = request.get("index_name")
    if index_name and index_name not in store.index_map:
        raise HTTPException(404, f"No such index: '{index_name}' exists.")
    ratio = request.get("context_token_ratio")
    if ratio and (ratio < 0.2 or ratio > 0.8):
        raise HTTPException(400, f"Invalid context_token_ratio: {ratio}. Must be between 0.2 and 0.8.")
    llm_params = {k: request[k] for k in ("model", "temperature", "top_p", "max_tokens") if request.get(k) is not None}
    if not isinstance(request.get("messages"), list) or not all(isinstance(m, dict) for m in request["messages"]):
        raise HTTPException(400, "Invalid request format: 'messages' must be a list of message objects")
    passthrough = {k: v for k, v in request.items() if k not in ("index_name", "context_token_ratio")}
    if not index_name:
        return llm.chat_completions_passthrough(passthrough)
    if request.get("tools") or request.get("functions"):
        return llm.chat_completions_passthrough(passthrough)

    for m in request["messages"]:
        role = m.get("role")
        if not role:
            raise HTTPException(400, "Invalid request format: messages must contain 'role'.")
        if role != "assistant" and m.get("content") is None:
            raise HTTPException(400, f"Invalid request format: messages must contain 'content' for role '{role}'.")
        if role not in ("user", "system", "assistant", "developer"):
            return llm.chat_completions_passthrough(passthrough)
        if role == "user":
            content = m.get("content")
            if not content:
                raise HTTPException(400, "Invalid request format: user messages must contain 'content'.")
            if isinstance(content, list):
                if any(not isinstance(p, str) and p.get("type") != "text" for p in content):
                    return llm.chat_completions_passthrough(passthrough)
            elif not isinstance(content, (str, dict)):
                return llm.chat_completions_passthrough(passthrough)

    max_tokens = request.get("max_tokens")
    total_prompt, history, user_parts, assistant_seen = "", [], [], False
    for m in reversed(request["messages"]):
        text = _message_text(m.get("content"))
        total_prompt = total_prompt + "\n\n" + text
        if m["role"] == "user" and not assistant_seen:
            user_parts.insert(0, text)          # the user messages after the last assistant turn form the query
        else:
            assistant_seen = assistant_seen or m["role"] == "assistant"
            history.insert(0, {"role": m["role"], "content": text})
    user_prompt = "\n\n".join(user_parts)
    if user_prompt == "":
        raise HTTPException(400, "There must be a user prompt since the latest assistant message.")
    prompt_len = llm.count_tokens(total_prompt)
    if prompt_len > llm.context_window:
        raise HTTPException(400, "Prompt length exceeds context window.")
    if max_tokens and max_tokens > llm.context_window - prompt_len:
        max_tokens = llm.context_window - prompt_len

    top_k = max(100, int((llm.context_window - prompt_len) / cfg["node_token_approximation"]))
    try:
        cands = store.dense_candidates(index_name, user_prompt, min(top_k, MAX_DENSE_K))
        chosen = select_context(cands, user_prompt, llm, ratio or cfg["context_token_fill_ratio"], max_tokens,
                                cfg["similarity_threshold"])
        if not chosen:
            return llm.chat_completions_passthrough(passthrough)            # base.py:404-412
        context_str = "\n\n".join(embed_text(n.text, n.metadata).strip() for n, _ in chosen)
        messages = [{"role": "system", "content": CONTEXT_TEMPLATE.format(context_str=context_str)}] + history + \
                   [{"role": "user", "content": user_prompt}]
        resp = llm.chat(messages, params=llm_params, max_tokens=max_tokens)
        answer = (resp.get("choices") or [{}])[0].get("message", {}).get("content", "")
        usage = llm.last_usage
        if not usage:
            p, c = llm.count_tokens(total_prompt), llm.count_tokens(answer or "")
            usage = {"prompt_tokens": p, "completion_tokens": c, "total_tokens": p + c}
        return {
            "id": uuid.uuid4().hex, "object": "chat.completion", "created": int(time.time()), "model": request.get("model"),
            "choices": [{"message": {"role": "assistant", "content": answer}, "finish_reason": "stop", "index": 0}],
            "source_nodes": [{"doc_id": n.ref_doc_id, "node_id": n.node_id, "text": n.text, "score": float(s),
                              "metadata": n.metadata} for n, s in chosen],
            "usage": usage,
        }
    except Exception as e:     # base.py:474-477 wraps everything raised in this block, HTTP errors of the LLM call included
        raise HTTPException(500, f"Chat completion failed: {e}")
