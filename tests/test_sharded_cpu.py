"""world_size-2 gloo test of the document-sharded /retrieve plumbing (kaito_b200.sharded): row
partition, global BM25 statistics all-reduce, all-gather layout, ordinal bases, merge + fuse.
The CUDA stages are replaced by oracle-backed stand-ins (this file is a test: it may use oracle/);
the result must equal the single-shard oracle pipeline over the whole corpus."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

PAD = np.uint64(0xFFFFFFFFFFFFFFFF)


def _obits(v):
    u = np.asarray(v, np.float32).view(np.uint32).astype(np.uint64)
    return np.where(u & np.uint64(0x80000000), ~u & np.uint64(0xFFFFFFFF), u | np.uint64(0x80000000))


def _from_obits(o):
    o = o.astype(np.uint64)
    u = np.where(o & np.uint64(0x80000000), o & np.uint64(0x7FFFFFFF), ~o & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    return u.view(np.float32)


def keys_asc(val, ordn):
    k = (_obits(val) << np.uint64(32)) | np.asarray(ordn, np.int64).astype(np.uint64)
    return np.where(np.asarray(ordn) < 0, PAD, k)


def keys_desc(val, ordn):
    k = ((~_obits(val) & np.uint64(0xFFFFFFFF)) << np.uint64(32)) | np.asarray(ordn, np.int64).astype(np.uint64)
    return np.where(np.asarray(ordn) < 0, PAD, k)


class OracleStages:
    """Test double for kaito_b200.sharded.NativeStages backed by the CPU oracle."""

    def __init__(self, o, x, csr, vocab):
        self.o, self.x, self.csr, self.vocab = o, x, csr, vocab
        self.post, self.base = None, 0

    def commit_local(self, vocab):
        off, ids, tf, dl = self.csr
        return self.o.bm25_df(off, ids, vocab), len(dl), int(dl.astype(np.int64).sum())

    def commit_global(self, vocab, df, n_docs, total_len, ordinal_base):
        off, ids, tf, dl = self.csr
        self.post = self.o.bm25_build(off, ids, tf, dl, vocab, df, n_docs, total_len)
        self.base = ordinal_base

    def dense_candidates(self, q, P, out):
        d, o = self.o.dense_topk(self.x, q.numpy()[:, : self.x.shape[1]], P)
        out.copy_(torch.from_numpy(keys_asc(d, np.where(o >= 0, o + self.base, -1)).view(np.int64)))

    def bm25_candidates(self, terms, toff, batch, P, out, toff_host=None):
        t, off = terms.numpy().view(np.uint32), toff.numpy()
        for b in range(batch):
            s, o = self.o.bm25_query(self.post, t[off[b]:off[b + 1]], P)
            out[b].copy_(torch.from_numpy(keys_desc(s, np.where(o >= 0, o + self.base, -1)).view(np.int64)))

    def merge(self, gathered, n_lists, batch, P, out):
        g = gathered.numpy().view(np.uint64).reshape(n_lists, batch, P)
        for b in range(batch):
            out[b].copy_(torch.from_numpy(np.sort(g[:, b, :].reshape(-1))[:P].view(np.int64)))

    def fuse(self, batch, P, k, dense_keys, bm25_keys, vw, tw, mode, out):
        dk = dense_keys.numpy().view(np.uint64)
        bk = None if bm25_keys is None else bm25_keys.numpy().view(np.uint64)
        for b in range(batch):
            dv = dk[b] != PAD
            dd, do = _from_obits(dk[b][dv] >> np.uint64(32)), (dk[b][dv] & np.uint64(0xFFFFFFFF)).astype(np.int64)
            if bk is None:
                n = min(k, len(do))
                out["ordinal"][b, :n] = torch.from_numpy(do[:n]); out["final"][b, :n] = torch.from_numpy(dd[:n].astype(np.float64))
                out["count"][b] = n
                continue
            bv = bk[b] != PAD
            bs = _from_obits(~(bk[b][bv] >> np.uint64(32)) & np.uint64(0xFFFFFFFF))
            bo = (bk[b][bv] & np.uint64(0xFFFFFFFF)).astype(np.int64)
            fin, de, sp, rk, od = self.o.fuse(dd, do, bs, bo, k, vw, tw, mode)
            n = len(od)
            out["final"][b, :n] = torch.from_numpy(fin); out["ordinal"][b, :n] = torch.from_numpy(od)
            out["dense"][b, :n] = torch.from_numpy(de); out["sparse"][b, :n] = torch.from_numpy(sp)
            out["rank"][b, :n] = torch.from_numpy(rk); out["count"][b] = n


def _worker(rank, world, port, n, d, vocab, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as o
        from kaito_b200.sharded import ShardedRetriever, shard_range
        x = o.synth_dense(n, d, 1)
        off, ids, tf, dl = o.synth_sparse(n, vocab, 2)
        q = o.synth_queries(x, 5, 3)
        qs = o.synth_query_terms(vocab, 5, 4, rank_offset=20)
        lo, hi = shard_range(n, world, rank)
        csr = (off[lo:hi + 1] - off[lo], ids[off[lo]:off[hi]], tf[off[lo]:off[hi]], dl[lo:hi])
        sr = ShardedRetriever(OracleStages(o, x[lo:hi], csr, vocab), torch.device("cpu"), d)
        n_docs, total, base = sr.commit(vocab, hi - lo)
        assert (n_docs, total, base) == (n, int(dl.astype(np.int64).sum()), lo)
        got = sr.retrieve(q, qs, k)
        got_dense_only = ShardedRetriever(OracleStages(o, x[lo:hi], csr, vocab), torch.device("cpu"), d)
        got_dense_only.stages.base = lo
        g2 = got_dense_only.retrieve(q, None, k)
        # single-shard oracle over the whole corpus
        post = o.bm25_build(off, ids, tf, dl, vocab)
        P = o.pool_size(k)
        for b in range(5):
            dd, do = o.dense_topk(x, q[b:b + 1], P)
            bs, bo = o.bm25_query(post, qs[b], P)
            fin, de, sp, rk, od = o.fuse(dd[0], do[0], bs, bo, k)
            c = int(got["count"][b])
            assert c == len(od)
            assert np.array_equal(got["ordinal"][b, :c], od) and np.array_equal(got["final"][b, :c], fin)
            assert np.array_equal(got["rank"][b, :c], rk)
            assert np.array_equal(g2["ordinal"][b, :k], do[0, :k])
        ret[rank] = "ok"
    except Exception as e:  # surface the failure in the parent
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("world,n", [(2, 2001), (2, 64), (3, 1000)])
def test_shards_equal_single_shard_oracle(world, n):
    """world_size 2 (odd and tiny corpora) and 3 (uneven shards, three-way gather/merge)"""
    from oracle import oracle as o
    o.build()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, 48, 300, 7, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret.get(r) == "ok", ret.get(r)


def test_shard_range_partition():
    from kaito_b200.sharded import shard_range
    for n in (0, 1, 7, 100, 12_500_001):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, w, i) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
