"""ctypes front-end of the CPU oracle (oracle/krag_oracle.c) plus small pure-Python
restatements used to cross-check the C code.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by kaito_b200/.

Parity status (see krag_oracle.c header): fusion PINNED against fixtures generated from
the reference's own hybrid_retriever.py; dense L2^2 / BM25 arithmetic PARITY UNPINNED
(faiss-cpu / bm25s are not vendored in /root/reference and not installable offline).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkrag_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc). Returns the .so path."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "krag_oracle.c"))
    ):
        subprocess.run(["make", "-C", _HERE, "-B", "libkrag_oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        f32p, i64p, u32p, u16p, f64p, i32p = (C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_uint32),
                                              C.POINTER(C.c_uint16), C.POINTER(C.c_double), C.POINTER(C.c_int32))
        L.krag_oracle_l2sq.argtypes = [f32p, C.c_int64, C.c_int32, f32p, f32p]
        L.krag_oracle_l2sq.restype = None
        L.krag_oracle_dense_topk.argtypes = [f32p, C.c_int64, C.c_int32, u32p, f32p, C.c_int32, C.c_int32, f32p, i64p]
        L.krag_oracle_dense_topk.restype = None
        L.krag_oracle_bm25_idf.argtypes = [u32p, C.c_int64, C.c_int64, f32p]
        L.krag_oracle_bm25_idf.restype = None
        L.krag_oracle_bm25_score.argtypes = [C.c_float, C.c_uint32, C.c_uint32, C.c_double]
        L.krag_oracle_bm25_score.restype = C.c_float
        L.krag_oracle_bm25_df.argtypes = [C.c_int64, i64p, u32p, C.c_int64, u32p]
        L.krag_oracle_bm25_df.restype = None
        L.krag_oracle_bm25_build_ext.argtypes = [C.c_int64, i64p, u32p, u16p, u32p, C.c_int64, u32p, C.c_int64,
                                                 C.c_int64, i64p, u32p, f32p]
        L.krag_oracle_bm25_build_ext.restype = None
        L.krag_oracle_bm25_query.argtypes = [C.c_int64, u32p, i64p, u32p, f32p, u32p, C.c_int32, C.c_int32,
                                             f32p, f32p, i64p]
        L.krag_oracle_bm25_query.restype = None
        L.krag_oracle_fuse.argtypes = [f32p, i64p, C.c_int32, f32p, i64p, C.c_int32, C.c_double, C.c_double,
                                       C.c_int32, C.c_int32, f64p, f32p, f32p, i32p, i64p]
        L.krag_oracle_fuse.restype = C.c_int32
        L.krag_oracle_threads.restype = C.c_int32
        L.krag_oracle_set_threads.argtypes = [C.c_int32]
        L.krag_oracle_set_threads.restype = None
        _lib = L
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def threads() -> int:
    return int(lib().krag_oracle_threads())


def set_threads(n: int) -> int:
    """OpenMP team size of the oracle (torchrun pins OMP_NUM_THREADS=1 for its children); returns the size in effect"""
    lib().krag_oracle_set_threads(int(n))
    return threads()


# ----------------------------------------------------------------------------- dense
def l2sq(x: np.ndarray, q: np.ndarray) -> np.ndarray:
    """Squared L2 of one query against every row (IndexFlatL2 metric; faiss_store.py:44)."""
    x = np.ascontiguousarray(x, np.float32)
    q = np.ascontiguousarray(q, np.float32).reshape(-1)
    out = np.empty(x.shape[0], np.float32)
    lib().krag_oracle_l2sq(_p(x, C.c_float), x.shape[0], x.shape[1], _p(q, C.c_float), _p(out, C.c_float))
    return out


def alive_bitmap(n: int, dead=()) -> np.ndarray:
    bm = np.zeros((n + 31) // 32, np.uint32)
    idx = np.setdiff1d(np.arange(n), np.asarray(list(dead), np.int64))
    np.bitwise_or.at(bm, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
    return bm


def dense_topk(x: np.ndarray, q: np.ndarray, P: int, alive: np.ndarray | None = None):
    """Top-P (L2^2 asc, ordinal asc). Returns (dist[nq,P] f32, ord[nq,P] i64, -1 padded)."""
    x = np.ascontiguousarray(x, np.float32)
    q = np.ascontiguousarray(q, np.float32).reshape(-1, x.shape[1])
    nq = q.shape[0]
    dist = np.empty((nq, P), np.float32)
    ordn = np.empty((nq, P), np.int64)
    lib().krag_oracle_dense_topk(_p(x, C.c_float), x.shape[0], x.shape[1], _p(alive, C.c_uint32),
                                 _p(q, C.c_float), nq, P, _p(dist, C.c_float), _p(ordn, C.c_int64))
    return dist, ordn


def np_l2sq_f64(x: np.ndarray, q: np.ndarray) -> np.ndarray:
    """fp64 ground truth for tolerance checks (|oracle - this| must be ~1e-6)."""
    d = x.astype(np.float64) - q.astype(np.float64).reshape(1, -1)
    return (d * d).sum(axis=1)


# ------------------------------------------------------------------------------ bm25
class Postings:
    """CSC postings (docs ascending inside a term) with precomputed fp32 scores."""

    def __init__(self, off, doc, score, n_docs):
        self.off, self.doc, self.score, self.n_docs = off, doc, score, n_docs


def bm25_df(term_offsets, term_ids, vocab: int) -> np.ndarray:
    term_offsets = np.ascontiguousarray(term_offsets, np.int64)
    term_ids = np.ascontiguousarray(term_ids, np.uint32)
    df = np.empty(vocab, np.uint32)
    lib().krag_oracle_bm25_df(len(term_offsets) - 1, _p(term_offsets, C.c_int64), _p(term_ids, C.c_uint32), vocab,
                              _p(df, C.c_uint32))
    return df


def bm25_build(term_offsets, term_ids, term_tf, doc_len, vocab: int, df_global=None, n_global=None,
               total_len_global=None) -> Postings:
    """bm25s-lucene score matrix (hybrid_retriever.py:122-125). Global stats optional (shards)."""
    term_offsets = np.ascontiguousarray(term_offsets, np.int64)
    term_ids = np.ascontiguousarray(term_ids, np.uint32)
    term_tf = np.ascontiguousarray(term_tf, np.uint16)
    doc_len = np.ascontiguousarray(doc_len, np.uint32)
    n = len(term_offsets) - 1
    if df_global is None:
        df_global = bm25_df(term_offsets, term_ids, vocab)
        n_global = n
        total_len_global = int(doc_len.astype(np.int64).sum())
    df_global = np.ascontiguousarray(df_global, np.uint32)
    nnz = int(term_offsets[-1])
    off = np.empty(vocab + 1, np.int64)
    doc = np.empty(nnz, np.uint32)
    score = np.empty(nnz, np.float32)
    lib().krag_oracle_bm25_build_ext(n, _p(term_offsets, C.c_int64), _p(term_ids, C.c_uint32), _p(term_tf, C.c_uint16),
                                     _p(doc_len, C.c_uint32), vocab, _p(df_global, C.c_uint32), int(n_global),
                                     int(total_len_global), _p(off, C.c_int64), _p(doc, C.c_uint32),
                                     _p(score, C.c_float))
    return Postings(off, doc, score, n)


def bm25_query(post: Postings, q_terms, P: int, alive: np.ndarray | None = None):
    """Top-P (score desc, ordinal asc); zero-score docs fill; ordinal -1 pads past n_live."""
    q_terms = np.ascontiguousarray(q_terms, np.uint32)
    acc = np.empty(post.n_docs, np.float32)
    score = np.empty(P, np.float32)
    ordn = np.empty(P, np.int64)
    lib().krag_oracle_bm25_query(post.n_docs, _p(alive, C.c_uint32), _p(post.off, C.c_int64), _p(post.doc, C.c_uint32),
                                 _p(post.score, C.c_float), _p(q_terms, C.c_uint32), len(q_terms), P,
                                 _p(acc, C.c_float), _p(score, C.c_float), _p(ordn, C.c_int64))
    return score, ordn


def py_bm25_score(df: int, n_docs: int, tf: int, dl: int, avgdl: float) -> np.float32:
    """Pure-Python restatement of one matrix entry (cross-checks the C code)."""
    idf = np.float32(math.log(1 + (n_docs - df + 0.5) / (df + 0.5)))
    k1, b = 1.5, 0.75
    tfc = tf / (k1 * ((1 - b) + b * dl / avgdl) + tf)
    return np.float32(float(idf) * tfc)


# ---------------------------------------------------------------------------- fusion
def pool_size(max_results: int, candidate_multiplier: float = 3.0) -> int:
    """HybridRetriever.__init__, hybrid_retriever.py:96-98."""
    return int(max_results * max(1.0, candidate_multiplier))


def fuse(dense_dist, dense_ord, bm25_score, bm25_ord, k: int, vector_weight=0.7, text_weight=0.3, mode: int = 0):
    dense_dist = np.ascontiguousarray(dense_dist, np.float32)
    dense_ord = np.ascontiguousarray(dense_ord, np.int64)
    bm25_score = np.ascontiguousarray(bm25_score, np.float32)
    bm25_ord = np.ascontiguousarray(bm25_ord, np.int64)
    fin = np.empty(k, np.float64)
    de = np.empty(k, np.float32)
    sp = np.empty(k, np.float32)
    rk = np.empty(k, np.int32)
    od = np.empty(k, np.int64)
    cnt = lib().krag_oracle_fuse(_p(dense_dist, C.c_float), _p(dense_ord, C.c_int64), len(dense_ord),
                                 _p(bm25_score, C.c_float), _p(bm25_ord, C.c_int64), len(bm25_ord),
                                 float(vector_weight), float(text_weight), k, mode,
                                 _p(fin, C.c_double), _p(de, C.c_float), _p(sp, C.c_float), _p(rk, C.c_int32),
                                 _p(od, C.c_int64))
    return fin[:cnt], de[:cnt], sp[:cnt], rk[:cnt], od[:cnt]


def py_fuse(vector_nodes, keyword_nodes, max_results, vector_weight=0.7, text_weight=0.3):
    """Line-by-line Python restatement of HybridRetriever._fuse (hybrid_retriever.py:132-166)
    over (id, score) tuples; ties broken by ascending id (reference: unspecified)."""
    total = vector_weight + text_weight
    w_v, w_t = vector_weight / total, text_weight / total
    vector_scores = {nid: (s if s is not None else 0.0) for nid, s in vector_nodes}
    keyword_ranks = {nid: idx for idx, (nid, _) in enumerate(keyword_nodes)}
    scored = []
    for nid in set(vector_scores) | set(keyword_ranks):
        vec = vector_scores.get(nid, 0.0)
        r = keyword_ranks.get(nid)
        text = 1.0 / (1.0 + r) if r is not None else 0.0
        scored.append((nid, w_v * vec + w_t * text))
    scored.sort(key=lambda t: (-t[1], t[0]))
    return scored[:max_results]


# --------------------------------------------------------------------------- doc ids
def generate_doc_id(text: str) -> str:
    """BaseVectorStore.generate_doc_id, vector_store/base.py:82-85."""
    return hashlib.sha256(text.encode("utf-8")).hexdigest()


# ------------------------------------------------------------- synthetic test inputs
def synth_dense(n: int, d: int, seed: int) -> np.ndarray:
    g = np.random.default_rng(seed)
    x = g.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True).astype(np.float32)
    return x


def synth_queries(x: np.ndarray, nq: int, seed: int, return_rows: bool = False):
    """Half perturbed corpus rows (planted neighbours), half random."""
    g = np.random.default_rng(seed)
    d = x.shape[1]
    q = g.standard_normal((nq, d), dtype=np.float32)
    rows = g.integers(0, x.shape[0], nq)
    half = nq // 2
    q[:half] = x[rows[:half]] + 0.1 * q[:half] / np.sqrt(d, dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True).astype(np.float32)
    q = np.ascontiguousarray(q, np.float32)
    return (q, rows[:half]) if return_rows else q


def synth_sparse(n: int, vocab: int, seed: int, mean_len: float = 96.0, zipf_s: float = 1.07):
    """Zipf term draws with log-normal doc lengths.
    Returns CSR (term_offsets i64[n+1], term_ids u32, term_tf u16, doc_len u32)."""
    g = np.random.default_rng(seed)
    dl = np.clip(g.lognormal(np.log(mean_len), 0.6, n), 8, 512).astype(np.int64)
    ranks = np.arange(1, vocab + 1, dtype=np.float64)
    cdf = np.cumsum(ranks ** (-zipf_s))
    cdf /= cdf[-1]
    tokens = np.searchsorted(cdf, g.random(int(dl.sum())))
    # per-document unique terms with counts, vectorised: sort (doc, term) pairs and run-length encode
    # (identical output to np.unique(tokens_of_doc, return_counts=True) document by document)
    doc_of = np.repeat(np.arange(n, dtype=np.int64), dl)
    key = np.sort(doc_of * np.int64(vocab) + tokens.astype(np.int64))
    first = np.concatenate([[True], key[1:] != key[:-1]])
    uk = key[first]
    cnt = np.diff(np.concatenate([np.nonzero(first)[0], [len(key)]]))
    offs = np.concatenate([[0], np.cumsum(np.bincount(uk // vocab, minlength=n))]).astype(np.int64)
    return (offs, (uk % vocab).astype(np.uint32), np.minimum(cnt, 65535).astype(np.uint16), dl.astype(np.uint32))


def synth_query_terms(vocab: int, nq: int, seed: int, zipf_s: float = 1.07, rank_offset: int = 100):
    g = np.random.default_rng(seed)
    ranks = np.arange(1 + rank_offset, vocab + 1, dtype=np.float64)
    cdf = np.cumsum(ranks ** (-zipf_s))
    cdf /= cdf[-1]
    out = []
    for _ in range(nq):
        m = int(g.integers(3, 9))
        out.append((np.searchsorted(cdf, g.random(m)) + rank_offset).astype(np.uint32))
    return out
