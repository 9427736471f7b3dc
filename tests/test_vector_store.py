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


# ------------------------------------------------------------------ round-1 advisor findings
def _sole_document_lifecycle(store):
    """index one doc, update it, delete it (the sequence of test/rage2e/rag_test.go), then reload an all-dead snapshot"""
    ids = store.index_documents("solo", [{"text": "the only document of this index", "metadata": {"v": 1}}])
    u = store.update_documents("solo", [{"doc_id": ids[0], "text": "the only document, second edition", "metadata": {"v": 2}}])
    assert [d["doc_id"] for d in u["updated_documents"]] == ids and not u["not_found_documents"]
    r = store.retrieve("solo", "second edition of the only document", 3)
    assert r["count"] == 1 and r["results"][0]["text"] == "the only document, second edition" and r["results"][0]["metadata"] == {"v": 2}
    listed = store.list_documents_in_index("solo")
    assert listed["total_items"] == 1 and listed["documents"][0]["text"].endswith("second edition")
    d = store.delete_documents("solo", ids)
    assert d == {"deleted_doc_ids": ids, "not_found_doc_ids": []}
    assert store.retrieve("solo", "anything at all", 3)["count"] == 0          # every row tombstoned: nothing comes back
    assert store.list_documents_in_index("solo")["total_items"] == 0
    return ids


def test_sole_document_update_delete_cpu(cpu_store, tmp_path):
    _sole_document_lifecycle(cpu_store)
    cpu_store.persist("solo", str(tmp_path / "s"))
    cpu_store.load("solo2", str(tmp_path / "s"))                 # a snapshot whose rows are all dead loads
    assert cpu_store.retrieve("solo2", "anything", 3)["count"] == 0


@pytest.mark.gpu
def test_sole_document_update_delete_gpu(ctx, tmp_path):
    store = VectorStore(HashingEmbedding(64), ctx)
    _sole_document_lifecycle(store)
    store.persist("solo", str(tmp_path / "s"))
    store.load("solo2", str(tmp_path / "s"))
    assert store.retrieve("solo2", "anything", 3)["count"] == 0
    again = store.index_documents("solo2", [{"text": "a fresh document after the purge"}])   # and the index keeps working
    assert [x["doc_id"] for x in store.retrieve("solo2", "fresh document", 3)["results"]] == again
    store.delete_index("solo"); store.delete_index("solo2")


class _FlakyEmbedding(HashingEmbedding):
    fail = False

    def get_text_embedding_batch(self, texts):
        if self.fail:
            raise RuntimeError("embedding backend unavailable")
        return super().get_text_embedding_batch(texts)


def _failed_insert_leaves_no_trace(store, emb):
    a = store.index_documents("tx", [{"text": "alpha doc about tensor cores"}])
    emb.fail = True
    with pytest.raises(RuntimeError):
        store.index_documents("tx", [{"text": "beta doc about shared memory"}])
    emb.fail = False
    assert store.list_documents_in_index("tx")["total_items"] == 1           # the failed document is not listed
    with pytest.raises(ValueError) as e:                                      # custom_transformer.py:39-41, raised before any state is touched
        store.index_documents("tx", [{"text": "delta doc"}, {"text": "def f(): pass", "metadata": {"split_type": "code"}}])
    assert "Language not specified" in str(e.value) and store.list_documents_in_index("tx")["total_items"] == 1
    b = store.index_documents("tx", [{"text": "beta doc about shared memory"}])     # the retry indexes it
    g = store.index_documents("tx", [{"text": "gamma doc about thread block clusters"}])
    assert b[0] in [x["doc_id"] for x in store.retrieve("tx", "shared memory beta", 3)["results"]]
    store.delete_documents("tx", b)                                           # deleting beta must not hit gamma's rows
    got = [x["doc_id"] for x in store.retrieve("tx", "thread block clusters gamma", 3)["results"]]
    assert g[0] in got and b[0] not in got and a[0] in got + a


def test_failed_insert_leaves_no_trace_cpu(oracle):
    from tests.oracle_engine import OracleEngine
    emb = _FlakyEmbedding(64)
    _failed_insert_leaves_no_trace(VectorStore(emb, OracleEngine(oracle)), emb)


@pytest.mark.gpu
def test_failed_insert_leaves_no_trace_gpu(ctx):
    emb = _FlakyEmbedding(64)
    store = VectorStore(emb, ctx)
    _failed_insert_leaves_no_trace(store, emb)
    store.delete_index("tx")


def test_load_errors_match_reference_and_keep_the_live_index(cpu_store, tmp_path):
    cpu_store.index_documents("live", DOCS)
    before = cpu_store.retrieve("live", "BM25 keyword retrieval", 3)
    with pytest.raises(HTTPException) as e:                                   # base.py:841-845
        cpu_store.load("live", str(tmp_path / "nowhere"), overwrite=True)
    assert e.value.status_code == 404 and e.value.detail == f"Path does not exist: {tmp_path / 'nowhere'}"
    bad = tmp_path / "corrupt"
    bad.mkdir()
    (bad / "docstore.json").write_text("{\"version\": 1, \"vocab\": [], \"ref_docs\": {}, \"nodes\": []}")   # engine file missing
    with pytest.raises(HTTPException) as e:
        cpu_store.load("live", str(bad), overwrite=True)
    assert e.value.status_code == 500 and e.value.detail.startswith("Loading failed:")
    assert cpu_store.retrieve("live", "BM25 keyword retrieval", 3) == before   # the served index survived the failed load


def test_filter_bitmap_cache_tracks_inserts(cpu_store):
    cpu_store.index_documents("f", DOCS)
    st = cpu_store.index_map["f"]
    bm1 = st.allow_bitmap({"type": "recipe"}).copy()
    assert bm1.tolist() == [1 << 2]
    cpu_store.index_documents("f", [{"text": "another recipe with garlic", "metadata": {"type": "recipe"}}])
    assert st.allow_bitmap({"type": "recipe"}).tolist() == [(1 << 2) | (1 << 5)]
    assert st.allow_bitmap({"type": "recipe"}) is st.allow_bitmap({"type": "recipe"})    # served from the cache


def test_rwlock_readers_share_writers_exclude():
    import threading, time
    from kaito_b200.vector_store import RWLock
    rw, log = RWLock(), []
    def reader(i):
        with rw.reader():
            log.append(("r+", i)); time.sleep(0.05); log.append(("r-", i))
    def writer():
        with rw.writer():
            with rw.writer():                      # re-entrant for the owning thread
                with rw.reader():
                    log.append(("w+", 0)); time.sleep(0.05); log.append(("w-", 0))
    ts = [threading.Thread(target=reader, args=(i,)) for i in range(3)] + [threading.Thread(target=writer)]
    for t in ts: t.start()
    for t in ts: t.join()
    wi = log.index(("w+", 0))
    assert log[wi + 1] == ("w-", 0)                                            # nothing interleaves with the writer
    assert log[:2] == [("r+", 0), ("r+", 1)] or log[0][0] == "w+"              # readers overlap each other


def test_code_documents_are_split_by_the_code_splitter(cpu_store):
    """split_type == "code" (custom_transformer.py:38-49): CodeSplitter(language) chunks (max_chars 1500), other documents go
    through SentenceSplitter(); both kinds stay retrievable and keep their doc id"""
    code = "import os\n\n\n" + "\n\n".join(f"def function_{i}(x):\n    value = x + {i}\n    return value * {i}\n" for i in range(120))
    ids = cpu_store.index_documents("code", [{"text": code, "metadata": {"split_type": "code", "language": "python"}},
                                             {"text": "plain prose document about retrieval"}])
    st = cpu_store.index_map["code"]
    code_nodes = [n for n in st.nodes if n.ref_doc_id == ids[0]]
    assert len(code_nodes) > 3 and all(len(n.text) <= 1500 for n in code_nodes)
    assert "".join(n.text for n in code_nodes).count("def function_") == 120          # nothing lost, nothing duplicated
    assert [n.text for n in st.nodes if n.ref_doc_id == ids[1]] == ["plain prose document about retrieval"]
    got = cpu_store.retrieve("code", "function_77 value", 5)
    assert ids[0] in [r["doc_id"] for r in got["results"]]


def test_sentence_splitter_token_budget_and_overlap():
    from kaito_b200.splitter import SentenceSplitter, split_sentences
    sp = SentenceSplitter(chunk_size=60, chunk_overlap=20)
    text = " ".join(f"Sentence {i} is short." for i in range(40)) + "\n\n\n" + "A second paragraph follows. It has two sentences."
    chunks = sp.split(text)
    assert len(chunks) > 2 and all(sp._count(c) <= 60 for c in chunks)
    assert "".join(split_sentences(text)) == text
    for a, b in zip(chunks, chunks[1:]):                       # consecutive chunks of a paragraph share a tail/head of whole sentences
        if "second paragraph" in b:
            continue
        tail = a.split(". ")[-1]
        assert tail.rstrip(".") in b
    assert sp.split("tiny") == ["tiny"]
    with pytest.raises(ValueError):                             # metadata longer than the chunk size (LlamaIndex raises the same)
        sp.split("text", "k: " + "v " * 200)
