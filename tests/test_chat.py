"""/v1/chat/completions (kaito_b200.chat) against the behaviour the reference's own API tests pin
(presets/ragengine/tests/api/test_chat_completions.py:46-137 basic RAG answer + metrics, :140-190 pass-through without index,
:193-264 tools, :461-474 unknown index, :477-496 / :1008-1028 invalid request, :499-556 role/content validation,
:1080-1120 prompt longer than the context window; tests/api/test_node_processing.py:44-170 context selection) and the
selection rule of contex_selection_node_processor.py:67-121.  The LLM is an httpx.MockTransport; the engine is the oracle
double on CPU and the CUDA engine under -m gpu."""
import json
import re

import httpx
import pytest
from starlette.testclient import TestClient

from kaito_b200 import chat
from kaito_b200.embedding import HashingEmbedding
from kaito_b200.service import create_app
from kaito_b200.vector_store import HTTPException, VectorStore

URL = "http://llm.test:5000/v1/chat/completions"
ANSWER = {"id": "chatcmpl-test123", "object": "chat.completion", "created": 1, "model": "mock-model",
          "choices": [{"index": 0, "message": {"role": "assistant", "content": "This is a helpful response about the test document."},
                       "finish_reason": "stop"}],
          "usage": {"prompt_tokens": 25, "completion_tokens": 12, "total_tokens": 37}}
DOCS = [{"text": "KAITO is a Kubernetes operator for AI workloads."}, {"text": "KAITO simplifies AI model deployment on Kubernetes."},
        {"text": "KAITO supports GPU provisioning for AI inference."}, {"text": "Chocolate chip cookies need butter and sugar."},
        {"text": "Pasta boiling requires salted water."}, {"text": "Rain clouds form when air cools rapidly.", "metadata": {"topic": "weather"}},
        {"text": "Soccer players need good ball control skills."}, {"text": "Books provide knowledge and entertainment."}]


class FakeLLM:
    def __init__(self, status=200, body=None):
        self.posts, self.status, self.body = [], status, body or ANSWER

    def __call__(self, request: httpx.Request):
        if request.url.path == "/v1/models":
            return httpx.Response(200, json={"data": [{"id": "mock-model", "max_model_len": 2048}]})
        self.posts.append(json.loads(request.content))
        assert request.headers["authorization"] == "Bearer s3cret"
        return httpx.Response(self.status, json=self.body)


def _client(engine, fake, threshold=1.9, window=64000):
    cfg = {"persist_dir": "storage", "llm_inference_url": URL, "similarity_threshold": threshold, "llm_context_window": window}
    llm = chat.LLMClient(URL, "s3cret", window, transport=httpx.MockTransport(fake))
    app = create_app(VectorStore(HashingEmbedding(64), engine), cfg, llm=llm)
    c = TestClient(app)
    assert c.post("/index", json={"index_name": "test_index", "documents": DOCS}).status_code == 200
    return c


def _exercise(engine):
    fake = FakeLLM()
    c = _client(engine, fake)
    # ---- RAG answer: context goes into one system message, the user prompt stays last, source nodes come back
    req = {"index_name": "test_index", "model": "mock-model", "temperature": 0.7, "max_tokens": 100, "context_token_ratio": 0.8,
           "messages": [{"role": "system", "content": "Be brief."}, {"role": "user", "content": "earlier question"},
                        {"role": "assistant", "content": "earlier answer"},
                        {"role": "user", "content": "What is the KAITO Kubernetes operator for AI workloads?"}]}
    r = c.post("/v1/chat/completions", json=req)
    assert r.status_code == 200, r.text
    body = r.json()
    assert body["object"] == "chat.completion" and body["model"] == "mock-model" and len(body["id"]) == 32
    assert body["choices"] == [{"message": {"role": "assistant", "content": ANSWER["choices"][0]["message"]["content"]},
                                "finish_reason": "stop", "index": 0}]
    assert body["usage"] == ANSWER["usage"]
    src = body["source_nodes"]
    assert len(src) > 0 and src[0]["text"] == DOCS[0]["text"] and len(src[0]["doc_id"]) == 64
    assert [s["score"] for s in src] == sorted(s["score"] for s in src) and all(s["score"] <= 1.9 for s in src)   # distances, nearest first
    sent = fake.posts[-1]
    assert sent["model"] == "mock-model" and sent["temperature"] == 0.7 and sent["max_tokens"] == 100
    roles = [m["role"] for m in sent["messages"]]
    assert roles == ["system", "system", "user", "assistant", "user"]
    assert sent["messages"][0]["content"].startswith("Use the context information below to assist the user.")
    assert DOCS[0]["text"] in sent["messages"][0]["content"]
    assert sent["messages"][-1]["content"] == "What is the KAITO Kubernetes operator for AI workloads?"
    m = c.get("/metrics").text
    assert len(re.findall(r'rag_chat_requests_total{status="success"} ([1-9]\d*).0', m)) == 1
    # metadata travels with the node text into the context (MetadataMode.LLM) and back in source_nodes
    r = c.post("/v1/chat/completions", json={"index_name": "test_index", "messages": [{"role": "user", "content": "Rain clouds form when air cools rapidly"}]})
    assert r.json()["source_nodes"][0]["metadata"] == {"topic": "weather"}
    assert "topic: weather\n\nRain clouds form" in fake.posts[-1]["messages"][0]["content"]

    # ---- pass-through: no index, tools, functions, unsupported role, non-text user content
    n = len(fake.posts)
    plain = {"model": "mock-model", "messages": [{"role": "user", "content": "hi"}]}
    r = c.post("/v1/chat/completions", json=plain)
    assert r.status_code == 200 and r.json()["source_nodes"] is None and r.json()["id"] == "chatcmpl-test123"
    assert fake.posts[-1] == plain
    for extra in ({"tools": [{"type": "function", "function": {"name": "f"}}]}, {"functions": [{"name": "f"}]}):
        r = c.post("/v1/chat/completions", json={"index_name": "test_index", **plain, **extra})
        assert r.status_code == 200 and r.json()["source_nodes"] is None
        assert "index_name" not in fake.posts[-1] and list(extra)[0] in fake.posts[-1]
    r = c.post("/v1/chat/completions", json={"index_name": "test_index", "messages": [{"role": "tool", "content": "x"}]})
    assert r.json()["source_nodes"] is None
    r = c.post("/v1/chat/completions", json={"index_name": "test_index", "messages": [
        {"role": "user", "content": [{"type": "image_url", "image_url": {"url": "http://x/y.png"}}]}]})
    assert r.json()["source_nodes"] is None
    assert len(fake.posts) == n + 5
    # list-of-text-parts user content is RAG material
    r = c.post("/v1/chat/completions", json={"index_name": "test_index", "messages": [
        {"role": "user", "content": [{"type": "text", "text": "KAITO supports GPU provisioning for AI inference."}]}]})
    assert r.json()["source_nodes"][0]["text"] == DOCS[2]["text"]

    # ---- validation (status + detail strings of base.py:190-342)
    r = c.post("/v1/chat/completions", json={"index_name": "nonexistent_index", "messages": [{"role": "user", "content": "q"}]})
    assert r.status_code == 404 and "No such index: 'nonexistent_index' exists" in r.json()["detail"]
    r = c.post("/v1/chat/completions", json={"model": "mock-model"})
    assert r.status_code == 400 and "Invalid request" in r.json()["detail"]
    r = c.post("/v1/chat/completions", json={"index_name": "test_index", "messages": [{"content": "q"}]})
    assert r.status_code == 400 and "messages must contain 'role'" in r.json()["detail"]
    r = c.post("/v1/chat/completions", json={"index_name": "test_index", "messages": [{"role": "user"}]})
    assert r.status_code == 400 and "messages must contain 'content' for role 'user'" in r.json()["detail"]
    r = c.post("/v1/chat/completions", json={"index_name": "test_index", "context_token_ratio": 0.9, "messages": [{"role": "user", "content": "q"}]})
    assert r.status_code == 400 and "Invalid context_token_ratio: 0.9" in r.json()["detail"]
    r = c.post("/v1/chat/completions", json={"index_name": "test_index", "messages": [
        {"role": "user", "content": "q"}, {"role": "assistant", "content": "a"}]})
    assert r.status_code == 400 and r.json()["detail"] == "There must be a user prompt since the latest assistant message."
    assert len(re.findall(r'rag_chat_requests_total{status="failure"} ([1-9]\d*).0', c.get("/metrics").text)) == 1


def test_chat_cpu(oracle):
    from tests.oracle_engine import OracleEngine
    _exercise(OracleEngine(oracle))


@pytest.mark.gpu
def test_chat_gpu(ctx):
    _exercise(ctx)


def test_chat_threshold_budget_and_errors(oracle):
    from tests.oracle_engine import OracleEngine
    # default threshold 0.85 on L2^2: an unrelated question leaves no context -> pass-through (base.py:404-412)
    fake = FakeLLM()
    c = _client(OracleEngine(oracle), fake, threshold=0.85)
    r = c.post("/v1/chat/completions", json={"index_name": "test_index", "messages": [{"role": "user", "content": "completely unrelated zebra"}]})
    assert r.status_code == 200 and r.json()["source_nodes"] is None
    r = c.post("/v1/chat/completions", json={"index_name": "test_index", "messages": [{"role": "user", "content": DOCS[4]["text"]}]})
    assert [s["text"] for s in r.json()["source_nodes"]] == [DOCS[4]["text"]] and r.json()["source_nodes"][0]["score"] < 1e-6
    # prompt longer than the window -> 400 (test_chat_completions.py:1080-1120)
    c2 = _client(OracleEngine(oracle), FakeLLM(), window=100)
    r = c2.post("/v1/chat/completions", json={"index_name": "test_index", "messages": [{"role": "user", "content": "This is a very long message. " * 50}]})
    assert r.status_code == 400 and "Prompt length exceeds context window" in r.json()["detail"]
    # LLM failure on the RAG path -> 500 "Chat completion failed: ..."; on the pass-through the LLM's status is kept
    c3 = _client(OracleEngine(oracle), FakeLLM(status=400, body={"error": "Invalid request"}))
    r = c3.post("/v1/chat/completions", json={"index_name": "test_index", "messages": [{"role": "user", "content": DOCS[0]["text"]}]})
    assert r.status_code == 500 and r.json()["detail"].startswith("Chat completion failed:")
    r = c3.post("/v1/chat/completions", json={"model": "mock-model", "messages": []})
    assert r.status_code == 400 and "Invalid request" in r.json()["detail"]
    # an endpoint that is not /chat/completions cannot take the pass-through (inference.py:274-279)
    llm = chat.LLMClient("http://llm.test:5000/v1/completions", transport=httpx.MockTransport(FakeLLM()))
    with pytest.raises(HTTPException) as e:
        llm.chat_completions_passthrough({"messages": []})
    assert e.value.status_code == 400 and "Chat completions not supported through endpoint" in e.value.detail


def test_select_context_rule():
    """contex_selection_node_processor.py:79-121: budget = int(min(max_tokens, window - query - 150) * ratio); nearest first;
    nodes over the threshold or over the remaining budget are skipped, later smaller ones still taken."""
    class N:
        def __init__(self, text):
            self.text = text

    llm = chat.LLMClient(None, context_window=1000)
    llm._encoder_failed = True                                   # len / 3 token approximation (inference.py:517-521)
    nodes = [(N("a" * 300), 0.30), (N("b" * 900), 0.10), (N("c" * 60), 0.50), (N("d" * 30), 0.90), (N("e" * 150), 0.20)]
    # budget = int(min(400, 1000 - 10 - 150) * 0.5) = 200 tokens; b (300 tok) does not fit, e (50), a (100), c (20) do; d is too far
    got = chat.select_context(nodes, "q" * 30, llm, 0.5, 400, 0.85)
    assert [n.text[0] for n, _ in got] == ["e", "a", "c"]
    assert chat.select_context(nodes, "q" * 30, llm, 0.5, None, None)[0][0].text[0] == "b"     # 420-token budget: b fits first
    assert chat.select_context(nodes, "q" * 3000, llm, 0.5, 400, 0.85) == []                     # nothing left after the query
    assert chat.select_context([], "q", llm, 0.5, 400, 0.85) == []


def test_chat_roles_max_tokens_and_usage_fallback(oracle):
    """test_chat_completions.py:559-623 system message, :719-776 developer role, :626-659 unsupported role = the LLM's own
    error through the pass-through, :1123-1250 max_tokens larger than what the window leaves is clamped (never a 422), and the
    usage block is estimated when the LLM returns none (base.py:428-446)."""
    from tests.oracle_engine import OracleEngine
    no_usage = {k: v for k, v in ANSWER.items() if k != "usage"}
    fake = FakeLLM(body=no_usage)
    c = _client(OracleEngine(oracle), fake, window=1000)
    q = DOCS[1]["text"]
    r = c.post("/v1/chat/completions", json={"index_name": "test_index", "model": "mock-model", "temperature": 0.5, "messages": [
        {"role": "system", "content": "You are a helpful AI assistant specializing in Kubernetes."},
        {"role": "developer", "content": "Answer in one sentence."}, {"role": "user", "content": q}], "max_tokens": 5000})
    assert r.status_code == 200, r.text
    body = r.json()
    assert len(body["source_nodes"]) > 0 and body["source_nodes"][0]["text"] == q
    sent = fake.posts[-1]
    assert [m["role"] for m in sent["messages"]] == ["system", "system", "developer", "user"]
    assert 0 < sent["max_tokens"] < 1000                               # clamped to what the 1000-token window leaves
    u = body["usage"]
    assert u["total_tokens"] == u["prompt_tokens"] + u["completion_tokens"] and u["completion_tokens"] > 0
    # unsupported role -> pass-through; the LLM's 400 comes back as 400 with its body in the detail
    bad = FakeLLM(status=400, body={"detail": "bad request format"})
    c2 = _client(OracleEngine(oracle), bad)
    r = c2.post("/v1/chat/completions", json={"model": "mock-model", "messages": [{"role": "function", "content": "Function response", "name": "f"}]})
    assert r.status_code == 400 and "bad request format" in r.json()["detail"]
    # assistant turns end the query: only user messages after the last assistant message are searched (base.py:311-330)
    fake3 = FakeLLM()
    c3 = _client(OracleEngine(oracle), fake3)
    r = c3.post("/v1/chat/completions", json={"index_name": "test_index", "messages": [
        {"role": "user", "content": DOCS[3]["text"]}, {"role": "assistant", "content": "noted"},
        {"role": "user", "content": "Pasta boiling requires"}, {"role": "user", "content": "salted water."}]})
    assert r.status_code == 200
    assert fake3.posts[-1]["messages"][-1] == {"role": "user", "content": "Pasta boiling requires\n\nsalted water."}
    assert [m["role"] for m in fake3.posts[-1]["messages"]] == ["system", "user", "assistant", "user"]
    assert r.json()["source_nodes"][0]["text"] == DOCS[4]["text"]


def test_select_context_against_reference_golden():
    """tests/golden/context_selection_reference.json was produced by executing the reference's own
    ContextSelectionProcessor._postprocess_nodes (oracle/gen_golden_context.py); our restatement must pick the same
    nodes in the same order for every recorded case (budget arithmetic, stable distance order, threshold, greedy skip)."""
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "context_selection_reference.json")
    fx = json.load(open(path))
    assert fx["meta"]["addition_prompt_tokens"] == chat.ADDITION_PROMPT_TOKENS

    class N:
        def __init__(self, text, nid):
            self.text, self.nid = text, nid

    assert len(fx["cases"]) >= 30 and sum(1 for c in fx["cases"] if c["selected"]) >= 15
    for c in fx["cases"]:
        llm = chat.LLMClient(None, context_window=c["window"])
        llm._encoder_failed = True                       # the fixture's count_tokens is the len/3 fallback
        nodes = [(N("t" * L, i), s) for i, (L, s) in enumerate(zip(c["text_lens"], c["scores"]))]
        got = chat.select_context(nodes, "q" * c["query_len"], llm, c["ratio"], c["max_tokens"], c["threshold"])
        assert [n.nid for n, _ in got] == c["selected"], c


def test_chat_llm_500_and_assistant_history(oracle):
    """test_chat_completions.py:779-813: the LLM answering 500 on the RAG path surfaces as 500 with "An unexpected error
    occurred" in the detail; :816-875: an assistant turn with content is history, the last user message is the query."""
    from tests.oracle_engine import OracleEngine
    c = _client(OracleEngine(oracle), FakeLLM(status=500, body={"error": "Internal server error"}))
    r = c.post("/v1/chat/completions", json={"index_name": "test_index", "model": "mock-model",
                                            "messages": [{"role": "user", "content": DOCS[6]["text"]}]})
    assert r.status_code == 500 and "An unexpected error occurred" in r.json()["detail"]
    fake = FakeLLM()
    c = _client(OracleEngine(oracle), fake)
    r = c.post("/v1/chat/completions", json={"index_name": "test_index", "model": "mock-model", "messages": [
        {"role": "user", "content": "Hello"}, {"role": "assistant", "content": "Hello! How can I help you?"},
        {"role": "user", "content": DOCS[7]["text"]}]})
    assert r.status_code == 200 and r.json()["choices"][0]["message"]["content"] == ANSWER["choices"][0]["message"]["content"]
    assert [m["role"] for m in fake.posts[-1]["messages"]] == ["system", "user", "assistant", "user"]
    assert fake.posts[-1]["messages"][-1]["content"] == DOCS[7]["text"]
