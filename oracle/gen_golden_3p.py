"""Pin the oracle's THIRD-PARTY arithmetic on the real wheels whenever they are importable.

The dense L2^2, the BM25 scores/ranks and the tokeniser/stemmer of the reference path live in pip wheels that are
neither vendored under /root/reference nor installable in this offline container:
    faiss-cpu==1.11.0                      IndexIDMap(IndexFlatL2).search          faiss_store.py:44-49
    bm25s (via llama-index-retrievers-bm25==0.6.5) + PyStemmer                     hybrid_retriever.py:122-125, :220
so oracle/krag_oracle.c restates their published algorithms and says "parity unpinned" for them.  This script is the
other half of that statement: run it anywhere those wheels exist (a developer machine, the reference's own image) and
it writes tests/golden/third_party_reference.json from the REAL libraries on fixed seeded inputs;
tests/test_oracle.py::test_third_party_golden then pins the restatement (and, through the GPU parity tests, the
kernels) on it.  Each section is generated only if its wheel imports; the file records which did and their versions.

    python oracle/gen_golden_3p.py          # writes tests/golden/third_party_reference.json (or reports what is missing)

Nothing here is imported by the product; nothing reads /root/reference.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "third_party_reference.json")

CORPUS = [
    "The quick brown fox jumps over the lazy dog near the riverbank.",
    "Retrieval augmented generation combines dense vector search with keyword matching.",
    "BM25 ranks documents by term frequency, inverse document frequency and length normalisation.",
    "GPU kernels stream the corpus from high bandwidth memory exactly once per batch of queries.",
    "Kubernetes operators reconcile custom resources into deployments and services.",
    "The cats were running quickly; running cats are happier than sleeping dogs!",
    "Stemming maps connected, connecting and connection to the same stem.",
    "An inverted index stores, for every term, the list of documents that contain it.",
    "Tensor cores multiply small matrix tiles; shared memory stages the operands.",
    "A document about nothing in particular, with numbers 12345 and under_scores and hyphen-ated words.",
]
QUERIES = ["running cats", "vector search keyword", "documents that contain the term", "matrix tiles in shared memory", "zebra"]


def dense_section():
    import faiss                                                                    # faiss-cpu
    g = np.random.default_rng(20260921)
    out = {"version": getattr(faiss, "__version__", "?"), "cases": []}
    for n, d, nq, k in [(1000, 64, 4, 10), (5000, 384, 3, 30), (300, 768, 2, 900)]:
        x = g.standard_normal((n, d)).astype(np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        q = g.standard_normal((nq, d)).astype(np.float32)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        ix = faiss.IndexIDMap(faiss.IndexFlatL2(d))                                 # faiss_store.py:44-49
        ix.add_with_ids(x, np.arange(n, dtype=np.int64))
        dist, ids = ix.search(q, min(k, n) if k <= n else k)
        out["cases"].append({"n": n, "d": d, "k": k, "seed_note": "default_rng(20260921), rows then queries, unit-normalised",
                             "x": x.tolist() if n * d <= 64_000 else None, "q": q.tolist(),
                             "dist": dist.tolist(), "ids": ids.tolist()})
    return out


def bm25_section():
    import bm25s
    import Stemmer                                                                  # PyStemmer
    stemmer = Stemmer.Stemmer("english")
    tok = bm25s.tokenize(CORPUS, stopwords="english", stemmer=stemmer, return_ids=True)   # BM25Retriever.from_defaults defaults
    retr = bm25s.BM25()                                                             # method="lucene", k1=1.5, b=0.75
    retr.index(tok)
    vocab = tok.vocab if hasattr(tok, "vocab") else tok[1]
    ids = tok.ids if hasattr(tok, "ids") else tok[0]
    out = {"version": getattr(bm25s, "__version__", "?"), "corpus_token_ids": [list(map(int, t)) for t in ids],
           "vocab": {str(k): int(v) for k, v in vocab.items()}, "queries": []}
    for qtext in QUERIES:
        qt = bm25s.tokenize([qtext], stopwords="english", stemmer=stemmer, return_ids=False)
        k = len(CORPUS)
        docs, scores = retr.retrieve(qt, k=k, corpus=None)
        out["queries"].append({"query": qtext, "query_tokens": list(qt[0]), "doc_ids": [int(i) for i in docs[0]],
                               "scores": [float(s) for s in scores[0]]})
    return out


def stem_section():
    import Stemmer
    st = Stemmer.Stemmer("english")
    words = sorted({w for text in CORPUS + QUERIES for w in __import__("re").findall(r"(?u)\b\w\w+\b", text.lower())})
    return {"words": words, "stems": st.stemWords(words)}


def main():
    doc = {"generated_by": "oracle/gen_golden_3p.py", "python": sys.version.split()[0], "sections": {}, "missing": {}}
    for name, fn in (("faiss_flat_l2", dense_section), ("bm25s_lucene", bm25_section), ("pystemmer_english", stem_section)):
        try:
            doc["sections"][name] = fn()
        except ImportError as e:
            doc["missing"][name] = str(e)
    if not doc["sections"]:
        print("none of faiss / bm25s / Stemmer import here: nothing written; missing =", doc["missing"])
        return 1
    with open(OUT, "w") as f:
        json.dump(doc, f)
    print("wrote", os.path.normpath(OUT), "sections:", sorted(doc["sections"]), "missing:", sorted(doc["missing"]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
