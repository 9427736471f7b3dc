"""Generate tests/golden/fuse_reference.json by RUNNING the reference's own
presets/ragengine/vector_store/retriever/hybrid_retriever.py (HybridRetriever.__init__,
_fuse, _retrieve) in this container.

llama_index is not installed offline, so the five names the module imports from it are
replaced by minimal stand-ins (plain containers; no logic of the path lives in them).
All arithmetic and control flow recorded in the fixture is executed by the unmodified
reference source read from /root/reference at generation time.  Run:

    python oracle/gen_golden.py            # needs /root/reference; writes tests/golden/

The fixtures travel with the repo; nothing reads /root/reference at test time.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference/presets/ragengine/vector_store/retriever/hybrid_retriever.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "fuse_reference.json")


def _install_stubs():
    class BaseRetriever:
        def __init__(self, *a, **k):
            pass

        def retrieve(self, q):
            return self._retrieve(q)

    class NodeWithScore:
        def __init__(self, node=None, score=None):
            self.node, self.score = node, score

    class _Plain:
        def __init__(self, *a, **k):
            self.args, self.kw = a, k

    class _Enum:
        EQ = "=="
        AND = "and"

    mods = {
        "llama_index": types.ModuleType("llama_index"),
        "llama_index.core": types.ModuleType("llama_index.core"),
        "llama_index.core.retrievers": types.ModuleType("llama_index.core.retrievers"),
        "llama_index.core.schema": types.ModuleType("llama_index.core.schema"),
        "llama_index.core.vector_stores": types.ModuleType("llama_index.core.vector_stores"),
        "llama_index.core.vector_stores.types": types.ModuleType("llama_index.core.vector_stores.types"),
    }
    mods["llama_index.core"].QueryBundle = _Plain
    mods["llama_index.core"].VectorStoreIndex = _Plain
    mods["llama_index.core.retrievers"].BaseRetriever = BaseRetriever
    mods["llama_index.core.schema"].NodeWithScore = NodeWithScore
    t = mods["llama_index.core.vector_stores.types"]
    t.FilterCondition, t.FilterOperator, t.MetadataFilter, t.MetadataFilters = _Enum, _Enum, _Plain, _Plain
    sys.modules.update(mods)
    return NodeWithScore


class _Node:
    def __init__(self, nid, metadata=None):
        self.node_id, self.metadata = nid, metadata or {}


class _FakeRetriever:
    def __init__(self, nodes):
        self._nodes = nodes

    def retrieve(self, q):
        return self._nodes


class _FakeIndex:
    """Stands in for VectorStoreIndex: as_retriever(similarity_top_k=P) returns the first P
    dense candidates, docstore.docs is non-empty."""

    def __init__(self, dense_nodes):
        self._dense = dense_nodes
        self.docstore = types.SimpleNamespace(docs={"x": 1})
        self.seen_top_k = None

    def as_retriever(self, similarity_top_k, filters=None):
        self.seen_top_k = similarity_top_k
        return _FakeRetriever(self._dense[:similarity_top_k])


def main():
    NodeWithScore = _install_stubs()
    spec = importlib.util.spec_from_file_location("ref_hybrid_retriever", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    rng = np.random.default_rng(20260921)
    cases = []
    # (a) pool size / weights as computed by the reference constructor
    init_cases = []
    for max_results, mult, vw, tw in [(10, 3.0, 0.7, 0.3), (5, 3.0, 0.7, 0.3), (300, 3.0, 0.7, 0.3), (1, 3.0, 0.7, 0.3),
                                      (7, 0.5, 0.7, 0.3), (7, 2.5, 1.0, 1.0), (33, 3.0, 0.2, 0.9)]:
        r = ref.HybridRetriever(index=None, max_results=max_results, candidate_multiplier=mult, vector_weight=vw,
                                text_weight=tw)
        init_cases.append({"max_results": max_results, "candidate_multiplier": mult, "vector_weight": vw,
                           "text_weight": tw, "pool": r._candidate_pool_size, "w_v": r._vector_weight,
                           "w_t": r._text_weight})

    # (b) _fuse on random candidate lists (float32 distances / scores as the wire carries)
    for ci in range(40):
        k = int(rng.choice([1, 3, 5, 10, 30, 100]))
        P = int(k * 3)
        n_d = int(rng.integers(0, P + 1)) if ci % 5 == 0 else P
        n_b = int(rng.integers(0, P + 1)) if ci % 7 == 0 else P
        universe = rng.permutation(4 * P + 8)
        d_ids = universe[:n_d]
        overlap = int(rng.integers(0, min(n_d, n_b) + 1))
        b_ids = np.concatenate([rng.permutation(d_ids)[:overlap], universe[n_d:n_d + (n_b - overlap)]])
        b_ids = rng.permutation(b_ids)
        d_dist = np.sort(rng.uniform(0.2, 1.9, n_d).astype(np.float32))
        b_score = -np.sort(-rng.uniform(0.0, 25.0, n_b).astype(np.float32))
        vw, tw = (0.7, 0.3) if ci % 4 else (float(rng.uniform(0.1, 1)), float(rng.uniform(0.1, 1)))
        r = ref.HybridRetriever(index=None, max_results=k, vector_weight=vw, text_weight=tw)
        vn = [NodeWithScore(node=_Node(int(i)), score=float(s)) for i, s in zip(d_ids, d_dist)]
        kn = [NodeWithScore(node=_Node(int(i)), score=float(s)) for i, s in zip(b_ids, b_score)]
        fused = r._fuse(vn, kn)
        cases.append({
            "k": k, "vector_weight": vw, "text_weight": tw,
            "dense_ids": [int(i) for i in d_ids], "dense_dist": [float(s) for s in d_dist],
            "bm25_ids": [int(i) for i in b_ids], "bm25_score": [float(s) for s in b_score],
            "out_ids": [int(n.node.node_id) for n in fused], "out_final": [float(n.score) for n in fused],
        })

    # (c) the whole _retrieve flow, including the keyword-side metadata post-filter (:227-235)
    flow = []
    for ci in range(8):
        k = int(rng.choice([2, 5, 10]))
        P = 3 * k
        ids = rng.permutation(6 * P)
        dense = [NodeWithScore(node=_Node(int(i), {"tag": "a" if i % 2 else "b"}), score=float(s))
                 for i, s in zip(ids[:P + 5], np.sort(rng.uniform(0.3, 1.8, P + 5).astype(np.float32)))]
        kw_ids = rng.permutation(ids[: 3 * P])[:P]
        kw = [NodeWithScore(node=_Node(int(i), {"tag": "a" if i % 2 else "b"}), score=float(s))
              for i, s in zip(kw_ids, -np.sort(-rng.uniform(0, 20, P).astype(np.float32)))]
        mf = {"tag": "a"} if ci % 2 else None
        idx = _FakeIndex(dense)
        r = ref.HybridRetriever(index=idx, max_results=k, metadata_filter=mf)
        r._build_bm25_retriever = lambda top_k, _kw=kw: _FakeRetriever(_kw[:top_k])
        out = r._retrieve("q")
        flow.append({
            "k": k, "metadata_filter": mf, "seen_top_k": idx.seen_top_k,
            "dense": [[int(n.node.node_id), float(n.score), n.node.metadata["tag"]] for n in dense],
            "keyword": [[int(n.node.node_id), float(n.score), n.node.metadata["tag"]] for n in kw],
            "out_ids": [int(n.node.node_id) for n in out], "out_final": [float(n.score) for n in out],
        })
    # BM25-unavailable fallback (:216-218): vector-only, cut to max_results
    idx = _FakeIndex(dense)
    r = ref.HybridRetriever(index=idx, max_results=4)
    r._build_bm25_retriever = lambda top_k: None
    out = r._retrieve("q")
    fallback = {"k": 4, "dense": [[int(n.node.node_id), float(n.score)] for n in dense],
                "out_ids": [int(n.node.node_id) for n in out], "out_score": [float(n.score) for n in out]}

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump({"source": "kaito-project/kaito presets/ragengine/vector_store/retriever/hybrid_retriever.py",
                   "init": init_cases, "fuse": cases, "retrieve_flow": flow, "vector_only_fallback": fallback}, f)
    print(f"wrote {OUT}: {len(init_cases)} init, {len(cases)} fuse, {len(flow)} flow cases")


if __name__ == "__main__":
    main()
