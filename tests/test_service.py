"""HTTP surface (kaito_b200.service) against the reference's API tests
(presets/ragengine/tests/api/test_main.py:28-62) and the e2e contract (test/rage2e/rag_test.go:1248-1384).
CPU: engine double backed by the oracle.  GPU (-m gpu): the CUDA engine."""
import re

import pytest
from starlette.testclient import TestClient

from kaito_b200.embedding import HashingEmbedding
from kaito_b200.service import create_app
from kaito_b200.vector_store import VectorStore

CFG = {"persist_dir": "storage", "llm_inference_url": None}


def _exercise(client, tmp_path):
    assert client.get("/health").json() == {"status": "Healthy", "detail": None}
    req = {"index_name": "test_index", "documents": [{"text": "This is a test document"}, {"text": "Another test document"}]}
    r = client.post("/index", json=req)
    assert r.status_code == 200
    doc1, doc2 = r.json()
    assert doc1["text"] == "This is a test document" and len(doc1["doc_id"]) == 64 and not doc1["metadata"]
    assert doc2["text"] == "Another test document" and len(doc2["doc_id"]) == 64
    m = client.get("/metrics")
    assert m.status_code == 200
    assert len(re.findall(r'rag_index_requests_total{status="success"} ([1-9]\d*).0', m.text)) == 1   # test_main.py:52-62
    # /retrieve: e2e sends an unknown field (rag_test.go:1254); top-1 doc id/text checks (:1343, :1363)
    r = client.post("/retrieve", json={"index_name": "test_index", "query": "another test document", "max_node_count": 2,
                                       "context_token_ratio": 0.5})
    assert r.status_code == 200
    body = r.json()
    assert body["query"] == "another test document" and body["count"] == len(body["results"]) <= 2
    assert {x["doc_id"] for x in body["results"]} <= {doc1["doc_id"], doc2["doc_id"]}
    for x in body["results"]:   # models.NodeWithScore: the reference serialises the three optional scores as null
        assert set(x) == {"doc_id", "node_id", "text", "score", "metadata", "dense_score", "sparse_score", "source"}
        assert x["dense_score"] is None and x["sparse_score"] is None and x["source"] is None
    assert client.post("/retrieve", json={"index_name": "nope", "query": "q"}).status_code == 404
    assert client.post("/retrieve", json={"index_name": "nope", "query": "q"}).json() == {"detail": "No such index: 'nope' exists."}
    assert client.post("/retrieve", json={"index_name": "test_index", "query": "  "}).json() == {"detail": "Query string cannot be empty."}
    assert client.post("/retrieve", json={"index_name": "test_index", "query": "q", "max_node_count": 0}).status_code == 422
    assert client.post("/retrieve", json={"index_name": "test_index", "query": "q", "max_node_count": 301}).status_code == 422
    assert client.get("/indexes").json() == ["test_index"]
    d = client.get("/indexes/test_index/documents", params={"limit": 1}).json()
    assert d["count"] == 1 and d["total_items"] == 2
    assert client.get("/indexes/test_index/documents", params={"metadata_filter": "{bad"}).status_code == 400
    r = client.post("/indexes/test_index/documents/delete", json={"doc_ids": [doc1["doc_id"], "x"]})
    assert r.json() == {"deleted_doc_ids": [doc1["doc_id"]], "not_found_doc_ids": ["x"]}
    assert client.post("/v1/chat/completions", json={"messages": []}).status_code == 503   # main.py:331-335
    m = client.get("/metrics").text
    for name in ("rag_indexes_retrieve_requests_total", "rag_indexes_retrieve_latency_seconds", "rag_retrieve_result_count",
                 "rag_vector_store_operation_latency_seconds", "rag_lowest_source_score", "rag_avg_source_score",
                 "e2e_request_latency_seconds", "num_requests_running", "rag_embedding_latency_seconds"):
        assert name in m, name
    return body


def test_service_cpu(oracle, tmp_path):
    from tests.oracle_engine import OracleEngine
    app = create_app(VectorStore(HashingEmbedding(64), OracleEngine(oracle)), dict(CFG))
    _exercise(TestClient(app), tmp_path)


@pytest.mark.gpu
def test_service_gpu(ctx, oracle, tmp_path):
    from tests.oracle_engine import OracleEngine
    store = VectorStore(HashingEmbedding(64), ctx)
    client = TestClient(create_app(store, dict(CFG)))
    a = _exercise(client, tmp_path)
    b = _exercise(TestClient(create_app(VectorStore(HashingEmbedding(64), OracleEngine(oracle)), dict(CFG))), tmp_path)
    assert a == b   # same ids, order and fp64 scores on the wire
    # persist / load round trip through the HTTP API (lifecycle hooks use these: lifecycle/manager.py:126-326)
    p = str(tmp_path / "snap")
    assert client.post("/persist/test_index", params={"path": p}).status_code == 200
    assert client.post("/load/test_index", params={"path": p}).status_code == 409
    assert client.post("/load/test_index", params={"path": p, "overwrite": "true"}).status_code == 200
    assert client.delete("/indexes/test_index").status_code == 200


def test_wire_models_match_reference_schemas():
    """tests/golden/wire_models_reference.json holds the JSON schemas of the reference's own pydantic models
    (presets/ragengine/models.py executed by oracle/gen_golden_models.py): same fields, required sets, types, defaults and
    limits on every request/response model of the routes this service keeps."""
    import json
    import os
    from kaito_b200 import service
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wire_models_reference.json")))["models"]
    keep = ("type", "default", "minimum", "maximum", "anyOf", "items", "$ref")
    assert len(fx) == 11
    for name, ref in fx.items():
        sch = getattr(service, name).model_json_schema()
        assert sorted(sch.get("required", [])) == ref["required"], name
        props = {k: {kk: vv for kk, vv in v.items() if kk in keep} for k, v in sch.get("properties", {}).items()}
        assert props == ref["properties"], (name, props, ref["properties"])


def test_prometheus_metrics_match_reference_registry(oracle):
    """tests/golden/prometheus_metrics_reference.json lists every metric the reference registers (name, kind, labels, buckets;
    produced by executing its prometheus_metrics.py): the service registers exactly that set, and the middleware labels
    e2e requests with success/failure on the tracked paths only (main.py:97-128)."""
    import json
    import os
    from prometheus_client.metrics import MetricWrapperBase
    from kaito_b200 import service
    from tests.oracle_engine import OracleEngine
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prometheus_metrics_reference.json")))
    app = create_app(VectorStore(HashingEmbedding(64), OracleEngine(oracle)), dict(CFG))
    mx = service.build_metrics(__import__("prometheus_client").CollectorRegistry())
    got = {}
    for name, m in mx.items():
        assert isinstance(m, MetricWrapperBase)
        kind = type(m).__name__
        ub = m._kwargs.get("buckets") if kind == "Histogram" else None
        got[m._name] = {"kind": kind, "labels": list(m._labelnames),
                        "buckets": None if kind != "Histogram" else [float(b) for b in (ub if ub is not None else m.DEFAULT_BUCKETS) if b != float("inf")]}
    assert got == fx["metrics"]
    c = TestClient(app)
    c.get("/health"); c.get("/metrics")                                   # untracked paths
    c.post("/retrieve", json={"index_name": "nope", "query": "q"})        # tracked; the handler returned (404) -> "success"
    text = c.get("/metrics").text
    assert 'e2e_request_total{status="success"} 1.0' in text and 'path=' not in text
    assert "rag_hybrid_top_k_requested_bucket" in text                    # registered like the reference's, never observed


def test_env_defaults_match_reference_config(monkeypatch):
    """tests/golden/config_reference.json = defaults of the reference's config.py (executed with the variables unset)."""
    import json
    import os
    from kaito_b200 import chat, service, vector_store
    d = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config_reference.json")))["defaults"]
    for k in list(os.environ):
        if k.startswith(("RAG_", "LLM_", "EMBEDDING_", "LOCAL_EMBEDDING", "VECTOR_DB", "DEFAULT_VECTOR_DB", "MODEL_ID")):
            monkeypatch.delenv(k)
    e, c = service.env_config(), chat.chat_config()
    assert (e["embedding_source"], e["embedding_model"], e["vector_db_type"], e["persist_dir"], e["llm_inference_url"]) == \
        (d["EMBEDDING_SOURCE_TYPE"], d["LOCAL_EMBEDDING_MODEL_ID"], d["VECTOR_DB_TYPE"], d["DEFAULT_VECTOR_DB_PERSIST_DIR"], d["LLM_INFERENCE_URL"])
    assert (c["llm_access_secret"], c["llm_context_window"], c["similarity_threshold"], c["context_token_fill_ratio"], c["node_token_approximation"]) == \
        (d["LLM_ACCESS_SECRET"], d["LLM_CONTEXT_WINDOW"], d["RAG_SIMILARITY_THRESHOLD"], d["RAG_DEFAULT_CONTEXT_TOKEN_FILL_RATIO"],
         d["RAG_DOCUMENT_NODE_TOKEN_APPROXIMATION"])
    assert vector_store.RAG_MAX_TOP_K == d["RAG_MAX_TOP_K"] == service.RAG_MAX_TOP_K


def test_remote_embedding_model(oracle):
    """embedding.remote of the CRD (embedding/remote_embedding.py:24-76): {"inputs": text} with a bearer token, the reply is
    the vector; the store indexes and retrieves with it, the embedding metrics carry mode="remote", failures surface."""
    import json
    import httpx
    import numpy as np
    from kaito_b200.embedding import RemoteEmbeddingModel
    from tests.oracle_engine import OracleEngine
    local = HashingEmbedding(32)
    seen = []

    def handler(request: httpx.Request):
        body = json.loads(request.content)
        seen.append((request.headers["authorization"], body))
        if body["inputs"] == "boom":
            return httpx.Response(503, json={"error": "down"})
        return httpx.Response(200, json=[float(x) for x in local.get_text_embedding(body["inputs"])])

    emb = RemoteEmbeddingModel("http://embedder/embedding", "tok3n", transport=httpx.MockTransport(handler))
    assert emb.get_embedding_dimension() == 32 and seen[0] == ("Bearer tok3n", {"inputs": "This is a dummy sentence."})
    with pytest.raises(RuntimeError, match="Failed to get embedding from remote model"):
        emb.get_text_embedding("boom")
    app = create_app(VectorStore(emb, OracleEngine(oracle)), {**CFG, "embedding_source": "remote"})
    c = TestClient(app)
    assert c.post("/index", json={"index_name": "r", "documents": [{"text": "alpha beta"}, {"text": "gamma delta"}]}).status_code == 200
    r = c.post("/retrieve", json={"index_name": "r", "query": "gamma delta", "max_node_count": 1}).json()
    assert r["count"] == 1
    assert np.allclose(emb.get_query_embedding("gamma delta"), local.get_text_embedding("gamma delta"))
    assert 'rag_embedding_requests_total{mode="remote",status="success"} 1.0' in c.get("/metrics").text


def test_retrieve_requests_are_coalesced(oracle):
    """concurrent POST /retrieve calls are answered by far fewer engine calls than requests (kaito_b200/batcher.py), each
    request gets exactly the answer the unbatched path gives, and per-request errors stay per request"""
    import threading
    from tests.oracle_engine import OracleEngine
    from kaito_b200.batcher import RetrieveBatcher
    store = VectorStore(HashingEmbedding(64), OracleEngine(oracle))
    docs = [{"text": f"document number {i} about topic {i % 7} and subject {i % 5}", "metadata": {"bucket": i % 3}} for i in range(60)]
    store.index_documents("c", docs)
    queries = [f"topic {i % 7} subject {i % 5}" for i in range(40)] + ["   "] * 2
    want = [store.retrieve("c", q, 4) if q.strip() else None for q in queries]
    b = RetrieveBatcher(store, max_batch=64, max_wait_s=0.05)
    got, errs = [None] * len(queries), [None] * len(queries)
    def go(i):
        try:
            got[i] = b.retrieve("c", queries[i], 4, None)
        except Exception as e:
            errs[i] = e
    ts = [threading.Thread(target=go, args=(i,)) for i in range(len(queries))]
    for t in ts: t.start()
    for t in ts: t.join()
    assert got[:40] == want[:40]
    assert all(e is not None and e.status_code == 400 for e in errs[40:]) and all(e is None for e in errs[:40])
    assert b.requests == len(queries) and b.batches <= 6 and b.max_seen >= 10, (b.batches, b.max_seen)
    # different (top_k, filter) groups inside one window go to separate engine calls, with the right answers
    f1, f2 = b.submit("c", queries[0], 2, None), b.submit("c", queries[0], 4, {"bucket": 1})
    assert f1.result() == store.retrieve("c", queries[0], 2) and f2.result() == store.retrieve("c", queries[0], 4, {"bucket": 1})
    with pytest.raises(Exception) as e:
        b.retrieve("missing", "q", 3, None)
    assert e.value.status_code == 404
    b.close()


def test_http_retrieve_goes_through_the_coalescer(oracle):
    from tests.oracle_engine import OracleEngine
    store = VectorStore(HashingEmbedding(64), OracleEngine(oracle))
    app = create_app(store, dict(CFG))
    client = TestClient(app)
    client.post("/index", json={"index_name": "h", "documents": [{"text": f"alpha beta {i}"} for i in range(10)]})
    r = client.post("/retrieve", json={"index_name": "h", "query": "alpha 3", "max_node_count": 3})
    want = store.retrieve("h", "alpha 3", 3)
    assert r.status_code == 200 and [(x["doc_id"], x["score"]) for x in r.json()["results"]] == [(x["doc_id"], x["score"]) for x in want["results"]]
    assert app.state.batcher.requests >= 1 and app.state.batcher.batches >= 1


def test_serialised_responses_equal_the_dict_path(oracle, monkeypatch):
    """VectorStore.retrieve_batch_bytes (cached per-node JSON fragments; what the HTTP fast path and the front-end workers send)
    parses to exactly retrieve_batch()'s dicts plus the null defaults of NodeWithScore -- unicode, metadata, blank queries, 404"""
    import json
    from tests.oracle_engine import OracleEngine
    from kaito_b200.batcher import RetrieveBatcher
    store = VectorStore(HashingEmbedding(64), OracleEngine(oracle))
    docs = [{"text": f"dokument número {i} über thema {i % 7} \"quoted\" \\ back\nnew", "metadata": {"bucket": i % 3, "tag": f"ü{i}"}} for i in range(30)]
    docs += [{"text": f"plain document {i} thema {i % 7}"} for i in range(10)]
    store.index_documents("s", docs)
    queries = [f"thema {i % 7} número" for i in range(12)] + ["  ", "ünï \"q\""]
    for flt in (None, {"bucket": 1}):
        want = store.retrieve_batch("s", queries, 5, flt)
        got = store.retrieve_batch_bytes("s", queries, 5, flt)
        for w, g in zip(want, got):
            if isinstance(w, Exception):
                assert isinstance(g, Exception) and g.status_code == w.status_code == 400
                continue
            body, count, scores = g
            for r in w["results"]:
                r.setdefault("dense_score", None); r.setdefault("sparse_score", None); r.setdefault("source", None)
            assert json.loads(body) == w and count == w["count"] and scores == [r["score"] for r in w["results"]]
    with pytest.raises(Exception) as e:
        store.retrieve_batch_bytes("missing", ["q"], 3, None)
    assert e.value.status_code == 404
    # through the coalescer: futures and sinks, both kinds in one window
    b = RetrieveBatcher(store, max_batch=64, max_wait_s=0.05)
    sunk = []
    f1 = b.submit("s", queries[0], 5, None)
    f2 = b.submit_bytes("s", queries[0], 5, None)
    assert b.submit_bytes("s", queries[1], 5, None, sink=sunk.append) is None
    b.submit_bytes("s", "  ", 5, None, sink=sunk.append)
    assert json.loads(f2.result()[0])["results"][0]["doc_id"] == f1.result()["results"][0]["doc_id"]
    import time
    t0 = time.time()
    while len(sunk) < 2 and time.time() - t0 < 5:
        time.sleep(0.01)
    assert len(sunk) == 2 and json.loads(sunk[0][0])["query"] == queries[1] and sunk[1].status_code == 400
    b.close()
    # KRAG_COMPONENT_SCORES=1 fills the optional fields: the bytes path follows
    monkeypatch.setenv("KRAG_COMPONENT_SCORES", "1")
    store2 = VectorStore(HashingEmbedding(64), OracleEngine(oracle))
    store2.index_documents("s", docs)
    w = store2.retrieve_batch("s", queries[:2], 3, None)
    g = store2.retrieve_batch_bytes("s", queries[:2], 3, None)
    for o in w:
        for r in o["results"]:
            r.setdefault("dense_score", None); r.setdefault("sparse_score", None); r.setdefault("source", None)
    assert [json.loads(x[0]) for x in g] == w


def test_coalescer_keeps_answers_apart_under_concurrency(oracle):
    """two dispatcher threads, futures and sinks mixed, 6 submitting threads: every request gets the answer to ITS query, grouped
    engine calls stay far below the request count, and the per-request 400s stay per request"""
    import json
    import threading
    from tests.oracle_engine import OracleEngine
    from kaito_b200.batcher import RetrieveBatcher
    store = VectorStore(HashingEmbedding(64), OracleEngine(oracle))
    store.index_documents("m", [{"text": f"document {i} about topic{i % 13} and subject{i % 5}"} for i in range(80)])
    b = RetrieveBatcher(store, max_batch=32, max_wait_s=0.003, dispatchers=2)
    n_threads, per = 6, 60
    bad = []
    done = threading.Semaphore(0)

    def work(t):
        pending = []
        for i in range(per):
            q = f"topic{(t * per + i) % 13} subject{i % 5} t{t}i{i}" if i % 17 else "   "
            if i % 3 == 0:
                box = []
                b.submit_bytes("m", q, 3, None, sink=box.append)
                pending.append((q, None, box))
            elif i % 3 == 1:
                pending.append((q, b.submit_bytes("m", q, 3, None), None))
            else:
                pending.append((q, b.submit("m", q, 3, None), None))
        for q, fut, box in pending:
            try:
                if box is not None:
                    import time
                    t0 = time.time()
                    while not box and time.time() - t0 < 20:
                        time.sleep(0.001)
                    out = box[0]
                    if isinstance(out, Exception):
                        raise out
                else:
                    out = fut.result(timeout=20)
                body = json.loads(out[0]) if isinstance(out, tuple) else out
                if body["query"] != q or body["count"] != len(body["results"]):
                    bad.append((q, body["query"]))
            except Exception as e:
                if not (q.strip() == "" and getattr(e, "status_code", None) == 400):
                    bad.append((q, repr(e)))
        done.release()

    ts = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    for t in ts: t.start()
    for t in ts: t.join(60)
    assert not bad, bad[:5]
    assert b.requests == n_threads * per and b.batches < b.requests // 3, (b.requests, b.batches)
    b.close()
