"""The request sequence of the reference's cluster e2e suite (test/rage2e/rag_test.go:318-352 with the curl bodies of
:981, :1169-1181, :1194-1205, :1218-1220, :1234, :1249-1255, :1404-1418, :1432, :1446) replayed against this service, checked
with the same `strings.Contains` expectations (ExpectedLogContent) and the field checks of createAndValidateRetrievalPod
(:1296-1375).  curl prints FastAPI's compact JSON, so the substring expectations are byte-level.
Embedding-model dependent parts (distance threshold of the chat context, the 0..1 score range of :1358-1361) are relaxed for
the deterministic hashing embedder and say so."""
import json

import httpx
from starlette.testclient import TestClient

from kaito_b200 import chat
from kaito_b200.embedding import HashingEmbedding
from kaito_b200.service import create_app
from kaito_b200.vector_store import VectorStore

TEXT = "Kaito is an operator that automates the AI/ML model inference or tuning workload in a Kubernetes cluster"


def test_rage2e_request_sequence(oracle, tmp_path):
    from tests.oracle_engine import OracleEngine
    llm_calls = []

    def llm(request: httpx.Request):
        if request.url.path == "/v1/models":
            return httpx.Response(200, json={"data": [{"id": "phi-3-mini-128k-instruct", "max_model_len": 131072}]})
        llm_calls.append(json.loads(request.content))
        return httpx.Response(200, json={"id": "cmpl-1", "object": "chat.completion", "created": 1, "model": "phi-3-mini-128k-instruct",
                                         "choices": [{"index": 0, "message": {"role": "assistant", "content": "KAITO is an operator."}, "finish_reason": "stop"}]})

    url = "http://workspace-svc/v1/chat/completions"
    cfg = {"persist_dir": str(tmp_path), "llm_inference_url": url, "similarity_threshold": 1.9}    # hashing embedder: L2^2 ~ 1.4 here
    client = chat.LLMClient(url, transport=httpx.MockTransport(llm))
    # max_tokens = 50 leaves int(50 * 0.5) = 25 tokens of context: the 17-word document fits with a BPE tokenizer (~22 tokens) but
    # not with the offline len/3 fallback (35), so the replay counts words
    client.count_tokens = lambda text: len(text.split())
    app = create_app(VectorStore(HashingEmbedding(64), OracleEngine(oracle)), cfg, llm=client)
    c = TestClient(app)

    # createAndValidateIndexPod (:1168-1192)
    r = c.post("/index", json={"index_name": "kaito", "documents": [{"text": TEXT, "metadata": {"author": "kaito", "category": "kaito"}}]})
    assert r.status_code == 200 and TEXT in r.text and r.text.startswith("[") and r.text.endswith("]")
    doc = r.json()[0]
    doc_id = doc["doc_id"]
    assert doc_id and doc["text"] == TEXT
    # verifyIndexExists (:980-992)
    assert '"kaito"' in c.get("/indexes").text
    # createAndValidateQueryChatMessagesPod (:1404-1418): no remote flag -> phi-3 model name; the answer carries the source text
    r = c.post("/v1/chat/completions", json={"index_name": "kaito", "model": "phi-3-mini-128k-instruct",
                                            "messages": [{"role": "user", "content": "what is kaito?"}], "max_tokens": 50, "temperature": 0})
    assert r.status_code == 200 and TEXT in r.text
    assert llm_calls[-1]["model"] == "phi-3-mini-128k-instruct" and llm_calls[-1]["temperature"] == 0 and llm_calls[-1]["max_tokens"] == 50
    # createAndValidateRetrievalPod (:1248-1384); the body carries a field the route does not know (context_token_ratio)
    r = c.post("/retrieve", json={"index_name": "kaito", "query": "What is KAITO?", "context_token_ratio": 0.5})
    assert r.status_code == 200 and '"query":' in r.text
    body = r.json()
    assert body["query"] == "What is KAITO?" and isinstance(body["results"], list) and len(body["results"]) > 0
    first = body["results"][0]
    assert isinstance(first["doc_id"], str) and isinstance(first["node_id"], str) and first["doc_id"] == doc_id and first["text"] == TEXT
    assert isinstance(first["score"], float) and first["score"] >= 0      # <= 1 holds when the model's L2^2 <= 1 (bge on related text)
    assert int(body["count"]) == len(body["results"])
    # createAndValidatePersistPod / LoadPod (:1431-1458): default path, `overwrite=True` as curl sends it
    r = c.post("/persist/kaito")
    assert r.status_code == 200 and "Successfully persisted index kaito" in r.text
    r = c.post("/load/kaito?overwrite=True")
    assert r.status_code == 200 and "Successfully loaded index kaito" in r.text
    assert c.post("/retrieve", json={"index_name": "kaito", "query": "What is KAITO?"}).json()["results"][0]["doc_id"] == doc_id
    # createAndValidateUpdateDocumentPod (:1193-1215)
    r = c.post("/indexes/kaito/documents", json={"documents": [{"doc_id": doc_id, "text": TEXT + ". It now has RAG capabilities.",
                                                                   "metadata": {"author": "kaito", "category": "ai-ml"}}]})
    assert r.status_code == 200 and '"updated_documents":[{"doc_id":"' + doc_id + '"' in r.text
    # createAndValidateDeleteDocumentPod (:1217-1231)
    r = c.post("/indexes/kaito/documents/delete", json={"doc_ids": [doc_id]})
    assert r.status_code == 200 and '"deleted_doc_ids":["' + doc_id + '"]' in r.text
    # createAndValidateDeleteIndexPod (:1233-1246)
    r = c.delete("/indexes/kaito")
    assert r.status_code == 200 and "Successfully deleted index kaito" in r.text
