"""The multi-GPU service engine (kaito_b200.sharded_engine) on CPU: two processes over gloo, oracle-backed stages.
Rank 0 runs the real VectorStore on a ShardedEngine, rank 1 the worker loop; every answer must equal the single-process
store over the single-shard oracle engine -- ids, order and fp64 scores -- through index / retrieve (with and without a
metadata filter, batched) / update / delete / persist / load."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

DOCS = [{"text": f"document {i} talks about topic{i % 11} and subject{i % 7} in area{i % 3}", "metadata": {"area": i % 3}} for i in range(90)]
QUERIES = ["topic3 subject2", "area1 document", "subject5 topic10 area2", "nothing matches this zebra", "document 17 talks"]


def _run(rank, world, port, tmp, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    try:
        from oracle import oracle as o
        from kaito_b200.embedding import HashingEmbedding
        from kaito_b200.sharded_engine import ShardedEngine, init_distributed
        from kaito_b200.vector_store import VectorStore
        from tests.oracle_engine import OracleEngine, OracleShardStages
        ctl = init_distributed(rank, world, rank, port, backend="gloo")
        eng = ShardedEngine(lambda name, dim, path=None: OracleShardStages(o, name, dim) if path is None else _load(o, name, dim, path),
                            torch.device("cpu"), None, ctl)
        if rank != 0:
            eng.serve()
            ret[rank] = "ok"
            return
        sharded, single = VectorStore(HashingEmbedding(48), eng), VectorStore(HashingEmbedding(48), OracleEngine(o))

        def same(**kw):
            for q in QUERIES:
                a, b = sharded.retrieve("ix", q, **kw), single.retrieve("ix", q, **kw)
                assert [(r["doc_id"], r["score"]) for r in a["results"]] == [(r["doc_id"], r["score"]) for r in b["results"]], (q, kw)
            ab = sharded.retrieve_batch("ix", QUERIES, kw.get("max_node_count", 5), kw.get("metadata_filter"))
            for q, r in zip(QUERIES, ab):
                assert r == sharded.retrieve("ix", q, **kw)

        for st in (sharded, single):
            st.index_documents("ix", DOCS[:50])
            st.index_documents("ix", DOCS[50:])                      # append: round-robin continues where it stopped
        same(max_node_count=5)
        same(max_node_count=12, metadata_filter={"area": 1})
        ids = [r["doc_id"] for r in single.retrieve("ix", QUERIES[0], 6)["results"]]
        for st in (sharded, single):
            st.delete_documents("ix", ids[:3])
            st.update_documents("ix", [{"doc_id": ids[3], "text": "a rewritten document about topic3 and subject2", "metadata": {"area": 9}}])
        same(max_node_count=7)
        assert [n.node_id for n, _ in sharded.dense_candidates("ix", QUERIES[2], 20)] == [n.node_id for n, _ in single.dense_candidates("ix", QUERIES[2], 20)]
        sharded.persist("ix", os.path.join(tmp, "snap"))
        sharded.load("ix2", os.path.join(tmp, "snap"))
        for q in QUERIES:
            assert sharded.retrieve("ix2", q, 5)["results"] == sharded.retrieve("ix", q, 5)["results"]
        sharded.delete_index("ix2")
        eng.shutdown()
        ret[rank] = "ok"
    except Exception:
        import traceback
        ret[rank] = traceback.format_exc()
        try:
            if rank == 0:
                eng.shutdown()
        except Exception:
            pass
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def _load(o, name, dim, path):
    from tests.oracle_engine import OracleEngine, OracleShardStages
    st = OracleShardStages(o, name, dim)
    st.ix = OracleEngine(o).load_index(name, path)
    st.ix.post = None
    return st


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_service_engine_equals_single_process(world, tmp_path):
    from oracle import oracle as o
    o.build()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_run, args=(world, _free_port(), str(tmp_path), ret), nprocs=world, join=True)
    for r in range(world):
        assert ret.get(r) == "ok", ret.get(r)
