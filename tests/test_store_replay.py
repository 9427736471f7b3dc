"""The reference's store-level test-suite (presets/ragengine/tests/vector_store/test_base_store.py:55-720 and
test_retrieve.py:54-145) replayed against kaito_b200.vector_store.VectorStore (engine double on CPU): same calls, same
assertions; documents are dicts here where the reference passes pydantic `Document` objects.  The code-splitting test
(:295-343, tree-sitter) is out of scope (501), embedding-model dependent expectations are marked."""
import os

import httpx
import pytest

from kaito_b200 import chat
from kaito_b200.embedding import HashingEmbedding
from kaito_b200.vector_store import HTTPException, VectorStore, generate_doc_id


@pytest.fixture
def store(oracle):
    from tests.oracle_engine import OracleEngine
    return VectorStore(HashingEmbedding(64), OracleEngine(oracle))


def D(text, **meta):
    return {"text": text, "metadata": meta or {"type": "text"}}


def test_index_documents(store):                                            # :55-68
    ids = store.index_documents("test_index", [D("First document"), D("Second document")])
    assert len(ids) == 2 and set(ids) == {generate_doc_id("First document"), generate_doc_id("Second document")}


def test_index_documents_isolation(store):                                  # :71-86 + faiss check_indexed_documents
    store.index_documents("index1", [D("First document in index1")])
    store.index_documents("index2", [D("First document in index2")])
    a = store.list_documents_in_index("index1")["documents"]
    b = store.list_documents_in_index("index2")["documents"]
    assert [d["text"] for d in a] == ["First document in index1"] and [d["text"] for d in b] == ["First document in index2"]


def test_add_document(store):                                               # :281-292
    store.index_documents("test_index", [D("Third document")])
    store.index_documents("test_index", [D("Fourth document")])
    assert store.document_exists("test_index", generate_doc_id("Fourth document"))


def _llm():
    def handler(request: httpx.Request):
        if request.url.path == "/v1/models":
            return httpx.Response(200, json={"data": [{"id": "mock-model", "max_model_len": 2048}]})
        return httpx.Response(200, json={"choices": [{"index": 0, "message": {"role": "assistant", "content": "This is the completion from the API"}}]})
    return chat.LLMClient("http://localhost:5000/v1/chat/completions", transport=httpx.MockTransport(handler))


def test_update_document(store):                                            # :346-417
    ids = store.index_documents("test_index", [D("Fifth document")])
    upd = {"doc_id": ids[0], "text": "Updated Fifth document", "metadata": {"type": "text"}}
    assert store.update_documents("test_index", [upd])["updated_documents"][0]["doc_id"] == ids[0]
    assert store.update_documents("test_index", [upd])["unchanged_documents"][0]["doc_id"] == ids[0]      # same text again
    assert store.document_exists("test_index", ids[0])                      # the id survives the update
    out = store.chat_completion({"index_name": "test_index", "model": "mock-model", "temperature": 0.7, "max_tokens": 100,
                                 "messages": [{"role": "user", "content": "What is the first document?"}]},
                                _llm(), {**chat.chat_config(), "similarity_threshold": 1.9})              # hashing-embedder distances
    assert out["source_nodes"] is not None and out["source_nodes"][0]["text"] == "Updated Fifth document"
    bad = {"doc_id": "baddocid", "text": "Updated Fifth document", "metadata": {"type": "text"}}
    assert store.update_documents("test_index", [bad])["not_found_documents"][0]["doc_id"] == "baddocid"


def test_delete_document(store):                                            # :422-443
    ids = store.index_documents("test_index", [D(f"Document {i}") for i in range(10)])
    res = store.delete_documents("test_index", ids)
    assert all(i in res["deleted_doc_ids"] for i in ids)
    assert store.delete_documents("test_index", ["baddocid"])["not_found_doc_ids"] == ["baddocid"]


def test_add_document_on_existing_index(store):                             # :446-465
    store.index_documents("test_add_index", [D("Initial Doc")])
    ids = store.index_documents("test_add_index", [D(f"Document {i}") for i in range(10)])
    resp = store.list_documents_in_index("test_add_index", limit=10, offset=1)
    assert all(doc["doc_id"] == ids[i] for i, doc in enumerate(resp["documents"])) and resp["total_items"] == 11


def test_persist_index(store, tmp_path):                                    # :468-472
    store.index_documents("test_index", [D("Test document")])
    store.persist("test_index", str(tmp_path / "storage"))
    assert os.path.exists(tmp_path / "storage")


def test_delete_index(store):                                               # :475-490: ten documents of 1 KiB of NUL bytes
    store.index_documents("test_index", [D((b"\x00" * 1024).decode()) for _ in range(10)])
    assert "test_index" in store.list_indexes()
    store.delete_index("test_index")
    assert "test_index" not in store.list_indexes()


def test_list_documents_in_index(store):                                    # :493-558
    store.index_documents("test_index", [D(f"Document {i}") for i in range(10)])
    L = lambda **kw: store.list_documents_in_index("test_index", **kw)      # noqa: E731
    r = L(limit=5, offset=0); assert len(r["documents"]) == 5 and r["total_items"] == 10
    r = L(limit=5, offset=5); assert len(r["documents"]) == 5 and r["total_items"] == 10
    r = L(limit=5, offset=10); assert len(r["documents"]) == 0 and r["total_items"] == 10
    r = L(limit=15, offset=0); assert len(r["documents"]) == 10 and r["total_items"] == 10
    r = L(limit=10, offset=0); assert len(r["documents"]) == 10 and r["total_items"] == 10
    r = L(limit=1, offset=0); assert len(r["documents"]) == 1 and r["total_items"] == 10
    assert len(L(limit=1, offset=0, max_text_length=5)["documents"][0]["text"]) == 5
    assert "Document" in L(limit=1, offset=0, max_text_length=None)["documents"][0]["text"]


def test_list_documents_with_filter_index(store):                           # :562-648
    store.index_documents("test_index", [D(f"Document {i}", type="text", filename=f"file_{i}", branch="main") for i in range(10)])
    L = lambda **kw: store.list_documents_in_index("test_index", **kw)      # noqa: E731
    r = L(limit=5, offset=0, metadata_filter={"filename": "file_1"})
    assert len(r["documents"]) == 1 and r["documents"][0]["metadata"]["filename"] == "file_1" and r["total_items"] == 1
    a = L(limit=5, offset=0, metadata_filter={"branch": "main"})
    b = L(limit=5, offset=5, metadata_filter={"branch": "main"})
    assert len(a["documents"]) == 5 and len(b["documents"]) == 5 and a != b and a["total_items"] == b["total_items"] == 10
    assert all(d["metadata"]["branch"] == "main" for d in a["documents"] + b["documents"])
    r = L(limit=5, offset=0, metadata_filter={"filename": "file_5", "branch": "main"})
    assert len(r["documents"]) == 1 and r["documents"][0]["metadata"]["filename"] == "file_5" and r["total_items"] == 1
    r = L(limit=5, offset=0, metadata_filter={"filename": "file_15", "branch": "main"})
    assert len(r["documents"]) == 0 and r["total_items"] == 0
    store.index_documents("test_index", [D(f"New Document {i}", type="text", filename=f"file_{i}", branch="new_branch") for i in range(7)])
    r = L(limit=1, offset=0, metadata_filter={"branch": "new_branch"})
    assert len(r["documents"]) == 1 and r["total_items"] == 7 and r["documents"][0]["metadata"]["branch"] == "new_branch"


def test_persist_and_load_as_separate_index(store, tmp_path):               # :651-720
    store.index_documents("test_index", [D(f"Document {i}", type="text", filename=f"file_{i}", branch="main") for i in range(10)])
    store.persist("test_index", str(tmp_path / "storage"))
    store.load("second_test_index", str(tmp_path / "storage"), overwrite=True)
    r = store.list_documents_in_index("second_test_index", limit=5, offset=0)
    assert len(r["documents"]) == 5 and r["total_items"] == 10
    store.delete_documents("second_test_index", [r["documents"][0]["doc_id"]])        # the copy does not touch the original
    first = store.list_documents_in_index("test_index", limit=10, offset=0)
    second = store.list_documents_in_index("second_test_index", limit=10, offset=0)
    assert len(first["documents"]) == 10 and len(second["documents"]) == 9
    d0 = second["documents"][0]
    upd = store.update_documents("second_test_index", [{"doc_id": d0["doc_id"], "text": "Modified text", "metadata": d0["metadata"]}])
    assert len(upd["updated_documents"]) == 1 and upd["updated_documents"][0]["text"] == "Modified text"
    dele = store.delete_documents("second_test_index", [d0["doc_id"]])
    assert dele["deleted_doc_ids"] == [d0["doc_id"]]


# ---- tests/vector_store/test_retrieve.py:54-145, literal documents and queries
def test_retrieve_basic(store):
    store.index_documents("test_index", [D("Python is a programming language", category="tech"),
                                         D("JavaScript is used for web development", category="tech"), D("The sky is blue", category="nature")])
    r = store.retrieve(index_name="test_index", query="What is Python?", max_node_count=3)
    assert r is not None and "query" in r and "results" in r and "count" in r
    assert r["query"] == "What is Python?" and r["count"] <= 3


def test_retrieve_max_node_count(store):
    store.index_documents("test_index", [D(f"Document {i}", index=i) for i in range(10)])
    assert store.retrieve(index_name="test_index", query="document", max_node_count=2)["count"] <= 2


def test_retrieve_default_max_node_count(store):
    store.index_documents("test_index", [{"text": f"Technology document {i}", "metadata": {}} for i in range(10)])
    assert store.retrieve(index_name="test_index", query="technology")["count"] <= 5


def test_retrieve_nonexistent_index(store):
    with pytest.raises(HTTPException) as e:
        store.retrieve(index_name="nonexistent_index", query="test query")
    assert e.value.status_code == 404


def test_retrieve_result_structure(store):
    store.index_documents("test_index", [D("Python is great", lang="python")])
    r = store.retrieve(index_name="test_index", query="Python programming", max_node_count=3)
    assert isinstance(r, dict) and "query" in r and "results" in r and "count" in r
    assert r["count"] > 0
    first = r["results"][0]
    for k in ("doc_id", "node_id", "text", "score", "metadata"):
        assert k in first
