"""Test double of kaito_b200._native.Context / Index backed by the CPU oracle, so the host-side
store logic (kaito_b200.vector_store) is exercised on CPU.  Test infrastructure only."""
import numpy as np


class OracleIndex:
    def __init__(self, o, name, dim):
        self.o, self.name, self.dim = o, name, dim
        self.x = np.zeros((0, dim), np.float32)
        self.ids = np.zeros(0, np.uint64)
        self.off = np.zeros(1, np.int64)
        self.tid = np.zeros(0, np.uint32); self.tf = np.zeros(0, np.uint16); self.dl = np.zeros(0, np.uint32)
        self.dead = set()
        self.post = None

    def add(self, node_ids, vecs, term_offsets=None, term_ids=None, term_tf=None, doc_len=None):
        self.x = np.concatenate([self.x, np.asarray(vecs, np.float32).reshape(-1, self.dim)])
        self.ids = np.concatenate([self.ids, np.asarray(node_ids, np.uint64)])
        self.off = np.concatenate([self.off, self.off[-1] + np.asarray(term_offsets, np.int64)[1:]])
        self.tid = np.concatenate([self.tid, term_ids]); self.tf = np.concatenate([self.tf, term_tf])
        self.dl = np.concatenate([self.dl, doc_len]); self.post = None

    def remove(self, node_ids):
        for i in node_ids:
            self.dead.add(int(np.nonzero(self.ids == i)[0][0]))
        return len(node_ids)

    def _alive(self):
        return self.o.alive_bitmap(len(self.ids), self.dead) if self.dead else None

    def commit(self, vocab):
        self.vocab = vocab
        o, n = self.o, len(self.ids)
        live = np.array([i not in self.dead for i in range(n)])
        df = np.zeros(vocab, np.uint32)
        for d in np.nonzero(live)[0]:
            df[self.tid[self.off[d]:self.off[d + 1]]] += 1
        # tombstoned docs keep their rows but lose their postings and do not count in N / avgdl
        keep = np.repeat(live, np.diff(self.off))
        off2 = np.concatenate([[0], np.cumsum(np.where(live, np.diff(self.off), 0))]).astype(np.int64)
        self.post = o.bm25_build(off2, self.tid[keep], self.tf[keep], self.dl, vocab, df, int(live.sum()),
                                 int(self.dl[live].astype(np.int64).sum()))

    def retrieve(self, q, q_terms_list, k, cand_mult=3.0, vector_weight=0.7, text_weight=0.3, fusion_mode=0,
                 keyword_allow_bitmap=None):
        o = self.o
        P = o.pool_size(k, cand_mult)
        B = len(q)
        out = {"final": np.zeros((B, k)), "ordinal": np.full((B, k), -1, np.int64), "count": np.zeros(B, np.int32)}
        alive = self._alive()
        if fusion_mode & 0x100 and keyword_allow_bitmap is not None:       # filter pushdown: eligible = allow & alive
            words = (len(self.ids) + 31) // 32
            alive = np.asarray(keyword_allow_bitmap[:words], np.uint32) & (alive if alive is not None else np.uint32(0xFFFFFFFF))
            keyword_allow_bitmap = None
        fusion_mode &= 0xFF
        for b in range(B):
            dd, do = o.dense_topk(self.x, q[b:b + 1], P, alive)
            if q_terms_list is None or self.post is None:
                n = min(k, int((do[0] >= 0).sum()))
                out["final"][b, :n] = dd[0, :n]; out["ordinal"][b, :n] = do[0, :n]; out["count"][b] = n
                continue
            bs, bo = o.bm25_query(self.post, q_terms_list[b], P, alive)
            if keyword_allow_bitmap is not None:
                keep = np.array([o_ >= 0 and (keyword_allow_bitmap[o_ >> 5] >> (o_ & 31)) & 1 for o_ in bo], bool)
                bs, bo = bs[keep], bo[keep]
            fin, de, sp, rk, od = o.fuse(dd[0], do[0], bs, bo, k, vector_weight, text_weight, fusion_mode)
            out["final"][b, :len(od)] = fin; out["ordinal"][b, :len(od)] = od; out["count"][b] = len(od)
        return out

    def search_dense(self, q, k):
        return self.o.dense_topk(self.x, np.asarray(q, np.float32).reshape(-1, self.dim), k, self._alive())

    def persist(self, path):
        import os
        np.savez(os.path.join(path, "oracle_index.npz"), x=self.x, ids=self.ids, off=self.off, tid=self.tid, tf=self.tf, dl=self.dl,
                 dead=np.array(sorted(self.dead), np.int64), vocab=np.int64(getattr(self, "vocab", 0)))

    def drop(self):
        pass


class OracleEngine:
    def __init__(self, o):
        self.o = o

    def create_index(self, name, dim):
        return OracleIndex(self.o, name, dim)

    def load_index(self, name, path):
        import os
        z = np.load(os.path.join(path, "oracle_index.npz"))
        ix = OracleIndex(self.o, name, z["x"].shape[1])
        ix.x, ix.ids, ix.off, ix.tid, ix.tf, ix.dl = z["x"], z["ids"], z["off"], z["tid"], z["tf"], z["dl"]
        ix.dead = set(int(d) for d in z["dead"])
        if int(z["vocab"]) > 0:
            ix.commit(int(z["vocab"]))
        return ix


# ------------------------------------------------------------------------------------------------------------------
# stage double for kaito_b200.sharded_engine (one shard of a round-robin sharded index), CPU + gloo
PAD = np.uint64(0xFFFFFFFFFFFFFFFF)


def _obits(v):
    u = np.asarray(v, np.float32).view(np.uint32).astype(np.uint64)
    return np.where(u & np.uint64(0x80000000), ~u & np.uint64(0xFFFFFFFF), u | np.uint64(0x80000000))


def _from_obits(o):
    o = o.astype(np.uint64)
    u = np.where(o & np.uint64(0x80000000), o & np.uint64(0x7FFFFFFF), ~o & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    return u.view(np.float32)


class OracleShardStages:
    """what kaito_b200.sharded.NativeStages is to the CUDA library, on the CPU oracle: local candidate lists as u64 keys,
    merge and fuse; plus the mutations the sharded engine issues"""

    def __init__(self, o, name, dim):
        self.o, self.ix = o, OracleIndex(o, name, dim)
        self.base, self.stride, self.dim = 0, 1, dim

    # mutations
    def add(self, node_ids, vecs, term_offsets=None, term_ids=None, term_tf=None, doc_len=None):
        self.ix.add(node_ids, vecs, term_offsets, term_ids, term_tf, doc_len)

    def remove(self, node_ids):
        return self.ix.remove(node_ids)

    def set_ordinal_map(self, base, stride):
        self.base, self.stride = base, stride

    def n_rows(self):
        return len(self.ix.ids)

    def dim_padded(self):
        return self.dim

    def persist(self, path):
        self.ix.persist(path)

    def drop(self):
        pass

    def _glob(self, rows):
        rows = np.asarray(rows, np.int64)
        return np.where(rows >= 0, self.base + rows * self.stride, -1)

    # commit
    def commit_local(self, vocab):
        ix, n = self.ix, len(self.ix.ids)
        live = np.array([i not in ix.dead for i in range(n)], bool)
        df = np.zeros(vocab, np.uint32)
        for d in np.nonzero(live)[0]:
            df[ix.tid[ix.off[d]:ix.off[d + 1]]] += 1
        return df, int(live.sum()), int(ix.dl[live].astype(np.int64).sum()) if n else 0

    def commit_global(self, vocab, df, n_docs, total_len, ordinal_base):
        ix, n = self.ix, len(self.ix.ids)
        self.base = ordinal_base
        live = np.array([i not in ix.dead for i in range(n)], bool)
        keep = np.repeat(live, np.diff(ix.off)) if n else np.zeros(0, bool)
        off2 = np.concatenate([[0], np.cumsum(np.where(live, np.diff(ix.off), 0))]).astype(np.int64)
        ix.post = self.o.bm25_build(off2, ix.tid[keep], ix.tf[keep], ix.dl, vocab, np.asarray(df, np.uint32), int(n_docs), int(total_len))

    # candidate stages
    def dense_candidates(self, q, P, out):
        import torch
        ix = self.ix
        if len(ix.ids) == 0:
            out.copy_(torch.from_numpy(np.full(tuple(out.shape), PAD, np.uint64).view(np.int64)))
            return
        d, o = self.o.dense_topk(ix.x, q.numpy()[:, : self.dim], P, ix._alive())
        keys = (_obits(d) << np.uint64(32)) | self._glob(o).astype(np.uint64)
        out.copy_(torch.from_numpy(np.where(o < 0, PAD, keys).view(np.int64)))

    def bm25_candidates(self, terms, toff, batch, P, out, toff_host=None):
        import torch
        t, off = terms.numpy().view(np.uint32), toff.numpy()
        for b in range(batch):
            s, o = self.o.bm25_query(self.ix.post, t[off[b]:off[b + 1]], P, self.ix._alive())
            keys = ((~_obits(s) & np.uint64(0xFFFFFFFF)) << np.uint64(32)) | self._glob(o).astype(np.uint64)
            out[b].copy_(torch.from_numpy(np.where(o < 0, PAD, keys).view(np.int64)))

    def merge(self, gathered, n_lists, batch, P, out):
        import torch
        g = gathered.numpy().view(np.uint64).reshape(n_lists, batch, P)
        for b in range(batch):
            out[b].copy_(torch.from_numpy(np.sort(g[:, b, :].reshape(-1))[:P].view(np.int64)))

    def fuse(self, batch, P, k, dense_keys, bm25_keys, vw, tw, mode, out, allow=None):
        import torch
        dk = dense_keys.numpy().view(np.uint64)
        bk = None if bm25_keys is None else bm25_keys.numpy().view(np.uint64)
        ab = None if allow is None else allow.numpy().view(np.uint32)
        out["count"].zero_()
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
            if ab is not None:
                keep = np.array([bool((ab[o_ >> 5] >> (o_ & 31)) & 1) for o_ in bo], bool)
                bs, bo = bs[keep], bo[keep]
            fin, de, sp, rk, od = self.o.fuse(dd, do, bs, bo, k, vw, tw, mode)
            n = len(od)
            out["final"][b, :n] = torch.from_numpy(fin); out["ordinal"][b, :n] = torch.from_numpy(od)
            out["dense"][b, :n] = torch.from_numpy(de); out["sparse"][b, :n] = torch.from_numpy(sp)
            out["rank"][b, :n] = torch.from_numpy(rk); out["count"][b] = n
