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
