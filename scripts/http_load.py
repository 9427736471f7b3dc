#!/usr/bin/env python
"""HTTP-level load test of POST /retrieve (the call a client of the RAGService makes): text queries in, JSON out, with
WordPiece + BM25 tokenisation, the K5 forward, the coalescer and response serialisation all on the clock.

    python scripts/http_load.py [--gpus G] [--docs N] [--seconds T] [--clients C] [--concurrency K]

Rank 0 runs the real service app (kaito_b200.service.create_app over VectorStore) under uvicorn in this process; with
--gpus G > 1 the engine is the ShardedEngine with G - 1 worker processes (exactly what service.main() starts).  The corpus
is synthetic (krag_synth_fill: N x 768 unit vectors + Zipf postings) because 10M documents cannot be pushed through /index in
a benchmark's time; the docstore is a lazy stand-in ("synthetic document <ordinal>").  Queries are strings of synthetic
terms ("t123 t4567 ..."); the embedder has bge-base shapes with random weights and a synthetic WordPiece vocabulary.
C client PROCESSES x K concurrent keep-alive connections each hammer the server for T seconds; the line reports completed
requests/s, latency percentiles, and the coalescer's batch statistics."""
import argparse
import asyncio
import json
import multiprocessing as mp
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VOCAB = 1 << 20


def synth_queries(n, seed):
    g = np.random.default_rng(seed)
    s, off = 1.07, 100
    lo, hi = float(off + 1) ** (1 - s), float(VOCAB + 1) ** (1 - s)
    out = []
    for _ in range(n):
        m = int(g.integers(3, 9))
        x = (lo + g.random(m) * (hi - lo)) ** (1.0 / (1 - s))
        out.append(" ".join(f"t{int(v)}" for v in np.clip(x.astype(np.int64) - 1, 0, VOCAB - 1)))
    return out


def client_proc(port, seconds, concurrency, seed, ret):
    import aiohttp

    async def run():
        qs = synth_queries(4096, seed)
        lat, done, stop = [], 0, time.perf_counter() + seconds
        conn = aiohttp.TCPConnector(limit=concurrency)
        async with aiohttp.ClientSession(connector=conn) as sess:
            async def worker(w):
                nonlocal done
                i = w
                while time.perf_counter() < stop:
                    body = {"index_name": "load", "query": qs[i % len(qs)], "max_node_count": 10}
                    t0 = time.perf_counter()
                    async with sess.post(f"http://127.0.0.1:{port}/retrieve", json=body) as r:
                        js = await r.json()
                        assert r.status == 200 and js["count"] == 10, (r.status, js)
                    lat.append(time.perf_counter() - t0)
                    done += 1
                    i += concurrency
            await asyncio.gather(*[worker(w) for w in range(concurrency)])
        return done, lat
    done, lat = asyncio.run(run())
    ret.put((done, lat[:: max(1, len(lat) // 2000)]))


def serve_and_load(a, app, store, check=True):
    import uvicorn
    fronts, rpc = [], None
    engine_port = a.port + 1 if a.http_workers > 0 else a.port
    srv = uvicorn.Server(uvicorn.Config(app, host="127.0.0.1", port=engine_port, log_level="warning", access_log=False))
    th = threading.Thread(target=srv.run, daemon=True)
    th.start()
    while not srv.started:
        time.sleep(0.05)
    if a.http_workers > 0:                     # what service.main() does with KRAG_HTTP_WORKERS=N
        import tempfile
        from kaito_b200 import frontend, vector_store as vs_
        from kaito_b200.rpc import RetrieveRpcServer
        from kaito_b200.service import RAG_MAX_TOP_K
        from fastapi import HTTPException as FHE
        rpc = RetrieveRpcServer(app.state.batcher, app.state.observe_retrieve, (vs_.HTTPException, FHE),
                                path=os.path.join(tempfile.gettempdir(), f"krag-rpc-{os.getpid()}.sock"))
        fronts = frontend.spawn(a.http_workers, "127.0.0.1", a.port, f"127.0.0.1:{engine_port}", rpc.path, None, RAG_MAX_TOP_K)
        import urllib.request
        for _ in range(200):                   # wait until a worker answers (proxied /health)
            try:
                urllib.request.urlopen(f"http://127.0.0.1:{a.port}/health", timeout=1).read()
                break
            except Exception:
                time.sleep(0.1)
        time.sleep(1.0)                        # ... and the slower ones are listening too
    if check:
        # correctness spot check at the HTTP level: the coalesced answer equals the direct single-query engine call
        import urllib.request
        for q in synth_queries(4, 1):
            req = urllib.request.Request(f"http://127.0.0.1:{a.port}/retrieve", json.dumps({"index_name": "load", "query": q, "max_node_count": 10}).encode(),
                                         {"Content-Type": "application/json"})
            got = json.loads(urllib.request.urlopen(req).read())
            want = store.retrieve("load", q, 10)
            assert [(r["doc_id"], r["score"]) for r in got["results"]] == [(r["doc_id"], r["score"]) for r in want["results"]]
    bt = app.state.batcher
    b0, r0 = bt.batches, bt.requests
    ret = mp.Queue()
    procs = [mp.Process(target=client_proc, args=(a.port, a.seconds, a.concurrency, 100 + i, ret)) for i in range(a.clients)]
    t0 = time.perf_counter()
    for p in procs:
        p.start()
    outs = [ret.get() for _ in procs]
    for p in procs:
        p.join()
    wall = time.perf_counter() - t0
    done = sum(o[0] for o in outs)
    lat = np.sort(np.concatenate([np.asarray(o[1]) for o in outs]))
    res = {"value": done / a.seconds, "unit": "requests/s",
           "latency_ms": {"p50": float(lat[len(lat) // 2] * 1e3), "p90": float(lat[int(len(lat) * 0.9)] * 1e3), "p99": float(lat[int(len(lat) * 0.99)] * 1e3)},
           "coalescer": {"engine_calls": bt.batches - b0, "requests": bt.requests - r0, "mean_batch": (bt.requests - r0) / max(1, bt.batches - b0),
                         "max_batch": bt.max_seen, "window_us": bt.max_wait_s * 1e6},
           "wall_s": wall}
    res["http_workers"] = a.http_workers
    for p in fronts:
        p.terminate()
    for p in fronts:
        p.wait(timeout=10)
    if rpc is not None:
        rpc.close()
    srv.should_exit = True
    th.join(timeout=5)
    bt.close()
    return res


def host_only(a):
    """--fake-engine: everything of the service except the GPU"""
    from kaito_b200 import vector_store as vs
    from kaito_b200.service import create_app
    from kaito_b200.text import WordPieceTokenizer
    sys.setswitchinterval(float(os.getenv("KRAG_GIL_SWITCH_S", "0.0002")))
    pieces = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + [f"t{i}" for i in range(20000)] + ["t"] + [str(d) for d in range(10)] + \
             [f"##{d}" for d in range(10)] + [f"##{d:02d}" for d in range(100)] + [f"##{d:03d}" for d in range(1000)]
    tok = WordPieceTokenizer(pieces)

    class Emb:
        def get_embedding_dimension(self):
            return a.dim

        def get_query_embedding_batch(self, queries):
            for q in queries:
                tok.encode(q)
            return np.zeros((len(queries), a.dim), np.float32)

    class Index:
        def retrieve(self, q, terms, k, **kw):
            time.sleep(a.fake_engine * 1e-3)                   # releases the GIL like the ctypes call into the engine
            B = len(q)
            o = (np.arange(B * k, dtype=np.int64).reshape(B, k) * 7919) % a.docs
            return {"final": np.linspace(1.0, 0.5, B * k).reshape(B, k), "ordinal": o, "count": np.full(B, k, np.int32)}

    class Nodes:
        def __len__(self):
            return a.docs

        def __getitem__(self, o):
            return vs._Node(f"n{int(o)}", f"doc{int(o)}", f"synthetic document {int(o)}", None, int(o))

    class Vocab:
        terms = []

        def __len__(self):
            return VOCAB

        def query_terms(self, text):
            return np.array([int(w[1:]) for w in text.split() if w[:1] == "t" and w[1:].isdigit()], np.uint32)

    store = vs.VectorStore(Emb(), None)
    st = vs._IndexState(Index())
    st.nodes, st.vocab, st.committed = Nodes(), Vocab(), True
    store.index_map["load"] = st
    for w in a.worker_counts:
        a.http_workers = w
        app = create_app(store, {"persist_dir": "storage", "llm_inference_url": None})
        res = serve_and_load(a, app, store, check=False)
        print(json.dumps({"metric": "http_retrieve_requests_per_sec", "n_gpus": 0, **res,
                          "config": {"workload": f"POST /retrieve, top-10, host only: engine call = sleep({a.fake_engine} ms)", "clients": a.clients,
                                     "concurrency_per_client": a.concurrency, "seconds": a.seconds}}), flush=True)
    os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--docs", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--seconds", type=float, default=8.0)
    ap.add_argument("--clients", type=int, default=8)
    ap.add_argument("--concurrency", type=int, default=64)
    ap.add_argument("--port", type=int, default=5077)
    ap.add_argument("--http-workers", default="0", help="front-end worker processes on the public port (KRAG_HTTP_WORKERS); a comma "
                    "list runs one load test per value against the same index, one JSON line each")
    ap.add_argument("--fake-engine", type=float, default=None, metavar="MS",
                    help="no GPU: the engine call sleeps MS milliseconds and returns arbitrary ordinals (the real tokenisers, "
                         "coalescer, docstore and JSON run) -- measures the ceiling of the Python host alone")
    a = ap.parse_args()
    a.worker_counts = [int(x) for x in str(a.http_workers).split(",")]
    a.http_workers = a.worker_counts[0]
    if a.fake_engine is not None:
        return host_only(a)

    import torch
    import uvicorn
    import bench
    from kaito_b200 import _native, sharded_engine as se, vector_store as vs
    from kaito_b200.embedding import GpuBertEmbedding
    from kaito_b200.service import create_app
    from kaito_b200.text import WordPieceTokenizer

    sys.setswitchinterval(float(os.getenv("KRAG_GIL_SWITCH_S", "0.0002")))     # as service.main() does
    ctx = _native.Context(device_id=0, rank=0, world_size=a.gpus)
    workers, eng = [], None
    if a.gpus > 1:
        import socket
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0)); mport = sk.getsockname()[1]
        workers = se.spawn_workers(a.gpus, mport)
        ctl = se.init_distributed(0, a.gpus, 0, mport)
        eng = se.ShardedEngine(se.native_stages_factory(ctx), torch.device("cuda", 0), None, ctl)
        index = eng.synth_index("load", a.dim, a.docs, VOCAB)
    else:
        index = ctx.create_index("load", a.dim)
        index.synth_fill(a.docs, row_base=0, seed=20260921, vocab=VOCAB)
        index.commit(VOCAB)
    # embedder: bge shapes, random weights, synthetic WordPiece vocabulary ("t<id>" tokens + digit pieces)
    name = {384: "bge-small", 768: "bge-base", 1024: "bge-large"}[a.dim]
    cfg = dict(bench.BGE[name], max_position_embeddings=512)
    pieces = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + [f"t{i}" for i in range(20000)] + ["t"] + [str(d) for d in range(10)] + \
             [f"##{d}" for d in range(10)] + [f"##{d:02d}" for d in range(100)] + [f"##{d:03d}" for d in range(1000)]
    pieces += [f"[unused{i}]" for i in range(cfg["vocab_size"] - len(pieces))]
    emb = GpuBertEmbedding(ctx, WordPieceTokenizer(pieces), cfg, bench.random_bert_state(bench.BGE[name]), query_instruction=None)

    class Nodes:                                   # lazy docstore of the synthetic corpus
        def __len__(self):
            return a.docs

        def __getitem__(self, o):
            return vs._Node(f"n{int(o)}", f"doc{int(o)}", f"synthetic document {int(o)}", None, int(o))

    class Vocab:                                   # "t<id>" -> term id
        terms = []

        def __len__(self):
            return VOCAB

        def query_terms(self, text):
            return np.array([int(w[1:]) for w in text.split() if w[:1] == "t" and w[1:].isdigit()], np.uint32)

    store = vs.VectorStore(emb, eng if eng is not None else ctx)
    st = vs._IndexState(index)
    st.nodes, st.vocab, st.committed = Nodes(), Vocab(), True
    store.index_map["load"] = st
    for w in a.worker_counts:
        a.http_workers = w
        app = create_app(store, {"persist_dir": "storage", "llm_inference_url": None})
        res = serve_and_load(a, app, store)
        print(json.dumps({
            "metric": "http_retrieve_requests_per_sec", "n_gpus": a.gpus, **res,
            "config": {"workload": f"POST /retrieve, top-10, {a.docs} docs x {a.dim} fp32 + BM25 postings (synthetic), text queries of 3-8 terms",
                       "clients": a.clients, "concurrency_per_client": a.concurrency, "seconds": a.seconds,
                       "on_the_clock": "HTTP parse, WordPiece + BM25 tokenisation, K5 forward, coalescer, dense+BM25+fuse, JSON response"},
            "spot_check": "4 queries: HTTP answer == direct engine call (ids and fp64 scores)",
        }), flush=True)
    if eng is not None:
        eng.shutdown()
        for w in workers:
            w.wait(timeout=30)
    os._exit(0)


if __name__ == "__main__":
    main()
