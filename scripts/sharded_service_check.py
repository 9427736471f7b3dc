"""torchrun --nproc-per-node G scripts/sharded_service_check.py : the multi-GPU SERVICE path on real GPUs.
Rank 0 builds the real VectorStore on a ShardedEngine (NativeStages on every rank, the other ranks run the worker loop) and
the same store on a single-GPU context; /retrieve answers -- ids, order, fp64 scores -- must be identical through
index / append / batched retrieve / metadata filter / delete / update / persist / load, at the HTTP level too."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from kaito_b200 import _native
from kaito_b200 import sharded_engine as se
from kaito_b200.embedding import HashingEmbedding
from kaito_b200.vector_store import VectorStore

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
ctl = se.init_distributed(rank, world, lr, int(os.environ["MASTER_PORT"]))
ctx = _native.Context(device_id=lr, rank=rank, world_size=world)
eng = se.ShardedEngine(se.native_stages_factory(ctx), torch.device("cuda", lr), None, ctl)
if rank != 0:
    eng.serve()
else:
    g = np.random.default_rng(0)
    words = [f"w{i}" for i in range(400)]
    docs = [{"text": " ".join(g.choice(words, 30)), "metadata": {"area": int(i % 4)}} for i in range(3000)]
    queries = [" ".join(g.choice(words, 4)) for _ in range(64)]
    single_ctx = _native.Context(device_id=lr)
    sharded, single = VectorStore(HashingEmbedding(96), eng), VectorStore(HashingEmbedding(96), single_ctx)
    for st in (sharded, single):
        st.index_documents("ix", docs[:2000])
        st.index_documents("ix", docs[2000:])

    def same(k, flt=None):
        a = sharded.retrieve_batch("ix", queries, k, flt)
        b = single.retrieve_batch("ix", queries, k, flt)
        for q, ra, rb in zip(queries, a, b):
            assert [(r["doc_id"], r["score"]) for r in ra["results"]] == [(r["doc_id"], r["score"]) for r in rb["results"]], (q, k, flt)
        assert sharded.retrieve("ix", queries[0], k, flt) == a[0]

    same(10); same(5, {"area": 2}); same(100)
    victims = [r["doc_id"] for r in single.retrieve("ix", queries[0], 8)["results"]]
    for st in (sharded, single):
        st.delete_documents("ix", victims[:4])
        st.update_documents("ix", [{"doc_id": victims[4], "text": "w1 w2 w3 rewritten " + queries[0], "metadata": {"area": 7}}])
    same(10)
    with tempfile.TemporaryDirectory() as tmp:
        sharded.persist("ix", tmp)
        sharded.load("ix2", tmp)
        a, b = sharded.retrieve_batch("ix2", queries, 10), sharded.retrieve_batch("ix", queries, 10)
        assert a == b
        sharded.delete_index("ix2")
    # HTTP level: concurrent requests through the coalescer
    from concurrent.futures import ThreadPoolExecutor
    from starlette.testclient import TestClient
    from kaito_b200.service import create_app
    app = create_app(sharded, {"persist_dir": "storage", "llm_inference_url": None})
    client = TestClient(app)
    def post(q):
        return client.post("/retrieve", json={"index_name": "ix", "query": q, "max_node_count": 10}).json()
    with ThreadPoolExecutor(32) as ex:
        got = list(ex.map(post, queries))
    want = single.retrieve_batch("ix", queries, 10)
    for ga, wb in zip(got, want):
        assert [(r["doc_id"], r["score"]) for r in ga["results"]] == [(r["doc_id"], r["score"]) for r in wb["results"]]
    print(f"sharded service check ok: world={world}, {len(queries)} queries, coalescer batches={app.state.batcher.batches} "
          f"max={app.state.batcher.max_seen}", flush=True)
    app.state.batcher.close()
    eng.shutdown()
    single_ctx.close()
for sh in list(eng.shards.values()):
    sh.stages.drop()
ctx.close()
if dist.is_initialized():
    dist.destroy_process_group()
