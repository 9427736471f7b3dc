"""Store-level tests mirroring the reference's own suites (presets/ragengine/tests/vector_store/
test_retrieve.py:54-145, test_base_store.py:64-68,153-157) against kaito_b200.vector_store.VectorStore.
Run on CPU against an oracle-backed engine double and, under -m gpu, against the CUDA engine: the two
must agree result for result (ids, order, scores)."""
import numpy as np
import pytest

from kaito_b200.embedding import HashingEmbedding
from kaito_b200.vector_store import HTTPException, VectorStore, generate_doc_id

DOCS = [
    {"text": "First document in the test index about retrieval engines", "metadata": {"type": "text"}},
    {"text": "Second document talks about GPU kernels and tensor cores", "metadata": {"type": "text"}},
    {"text": "Third document is a cooking recipe with tomatoes and basil", "metadata": {"type": "recipe"}},
    {"text": "Fourth document describes BM25 keyword retrieval and inverted indexes", "metadata": {"type": "text"}},
    {"text": "Fifth document: vector databases store embeddings for similarity search"},
]


def _cpu_store(oracle):
    from tests.oracle_engine import OracleEngine
    return VectorStore(HashingEmbedding(64), OracleEngine(oracle))


@pytest.fixture
def cpu_store(oracle):
    return _cpu_store(oracle)


def _check_suite(store):
    ids = store.index_documents("test_index", DOCS)
    assert ids == [generate_doc_id(d["text"]) for d in DOCS] and all(len(i) == 64 for i in ids)
    # structure (test_retrieve.py:54-87)
    r = store.retrieve("test_index", "What is the first document?", max_node_count=5)
    assert r["query"] == "What is the first document?" and r["count"] == len(r["results"]) <= 5
    for res in r["results"]:
        assert set(res) == {"doc_id", "node_id", "text", "score", "metadata"}
        assert res["doc_id"] in ids
    # count limit (test_retrieve.py:104-124)
    assert store.retrieve("test_index", "document", max_node_count=2)["count"] <= 2
    # errors (test_retrieve.py:127-145, base.py:884-896)
    with pytest.raises(HTTPException) as e:
        store.retrieve("missing", "q")
    assert e.value.status_code == 404 and e.value.detail == "No such index: 'missing' exists."
    with pytest.raises(HTTPException) as e:
        store.retrieve("test_index", "   ")
    assert e.value.status_code == 400 and e.value.detail == "Query string cannot be empty."
    # append dedupes by doc id (base.py:99-131)
    again = store.index_documents("test_index", DOCS[:2] + [{"text": "Sixth document about Kubernetes operators"}])
    assert again[:2] == ids[:2] and store.list_documents_in_index("test_index", limit=100)["total_items"] == 6
    # keyword-side metadata filter (hybrid_retriever.py:227-235)
    r_f = store.retrieve("test_index", "document retrieval", max_node_count=3, metadata_filter={"type": "recipe"})
    assert r_f["count"] <= 3
    # filter pushdown (KRAG_FILTER_PUSHDOWN, not the reference): the bitmap restricts both scans, so every result matches
    store.filter_pushdown = True
    r_p = store.retrieve("test_index", "document retrieval", max_node_count=3, metadata_filter={"type": "recipe"})
    n_match = sum(1 for d in DOCS if (d.get("metadata") or {}).get("type") == "recipe")
    assert r_p["count"] == min(3, n_match) and all((x["metadata"] or {}).get("type") == "recipe" for x in r_p["results"])
    assert store.retrieve("test_index", "document retrieval", max_node_count=3, metadata_filter={"type": "nothing"})["count"] == 0
    store.filter_pushdown = False
    # delete
    d = store.delete_documents("test_index", [ids[0], "nope"])
    assert d == {"deleted_doc_ids": [ids[0]], "not_found_doc_ids": ["nope"]}
    r2 = store.retrieve("test_index", "What is the first document?", max_node_count=5)
    assert ids[0] not in [x["doc_id"] for x in r2["results"]]
    return r, r_f, r2, r_p


def test_store_suite_cpu(cpu_store):
    _check_suite(cpu_store)


def test_similarity_mode_orders_first_document_first(oracle):
    """test_base_store.py:153-157 asserts the top source node for "What is the first document?"; in the
    reference that ordering comes from the chat path (distance ascending).  The dense list the store feeds
    the fusion has the matching document first."""
    from tests.oracle_engine import OracleEngine
    eng = OracleEngine(oracle)
    store = VectorStore(HashingEmbedding(64), eng)
    store.index_documents("t", DOCS)
    st = store.index_map["t"]
    q = np.asarray(store.embed_model.get_query_embedding("first document test index retrieval engines"), np.float32)
    out = st.index.retrieve(q.reshape(1, -1), None, 1)
    assert st.nodes[int(out["ordinal"][0, 0])].text.startswith("First document")


@pytest.mark.gpu
def test_store_suite_gpu_equals_cpu_double(ctx, oracle):
    gpu_store = VectorStore(HashingEmbedding(64), ctx)
    a = _check_suite(gpu_store)
    b = _check_suite(_cpu_store(oracle))
    for ra, rb in zip(a, b):
        assert [x["doc_id"] for x in ra["results"]] == [x["doc_id"] for x in rb["results"]]
        assert [x["score"] for x in ra["results"]] == [x["score"] for x in rb["results"]]
    gpu_store.delete_index("test_index")


@pytest.mark.gpu
def test_store_persist_load_gpu(ctx, tmp_path):
    store = VectorStore(HashingEmbedding(64), ctx)
    store.index_documents("p", DOCS)
    a = store.retrieve("p", "keyword retrieval with BM25", 4)
    store.persist("p", str(tmp_path / "p"))
    with pytest.raises(HTTPException) as e:
        store.load("p", str(tmp_path / "p"))
    assert e.value.status_code == 409
    store.load("p2", str(tmp_path / "p"))
    b = store.retrieve("p2", "keyword retrieval with BM25", 4)
    assert a["results"] == b["results"]
    store.delete_index("p"); store.delete_index("p2")
