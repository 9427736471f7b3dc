"""bench.py's in-bench correctness check has its own host-side merge / fuse (numpy, independent of the library under test): pin those
helpers on the oracle, so a green `check` in a bench line means what it says."""
import numpy as np

import bench


def test_host_fuse_equals_the_oracle_fuse(oracle):
    g = np.random.default_rng(5)
    for trial in range(40):
        P, k = int(g.integers(3, 40)), int(g.integers(1, 12))
        do = g.choice(200, P, replace=False).astype(np.int64)
        dd = np.sort(g.random(P).astype(np.float32) * 2)
        nb = int(g.integers(0, P + 1))
        bo = g.choice(200, nb, replace=False).astype(np.int64)
        bs = -np.sort(-g.random(nb).astype(np.float32) * 9)
        if trial % 5 == 0 and P > 4:                      # padded tails, as the candidate lists carry them
            do[-2:] = -1
        fin, de, sp, rk, od = oracle.fuse(dd, do, bs, bo, k)
        got = bench._host_fuse(dd, do, bs, bo, k)
        assert [o for _, o in got] == od.tolist()
        assert [f for f, _ in got] == fin.tolist()


def test_key_values_inverts_the_key_layout():
    g = np.random.default_rng(2)
    vals = np.concatenate([g.standard_normal(50).astype(np.float32) * 3, np.array([0.0, 1e-30, 7.5], np.float32)])
    ords = g.integers(0, 2**31 - 1, vals.size).astype(np.uint64)

    def ordered(v):
        u = v.view(np.uint32).astype(np.uint64)
        return np.where(u & np.uint64(0x80000000), ~u & np.uint64(0xFFFFFFFF), u | np.uint64(0x80000000))
    asc = (ordered(vals) << np.uint64(32)) | ords                                             # dense keys: value ascending
    desc = ((~ordered(vals) & np.uint64(0xFFFFFFFF)) << np.uint64(32)) | ords                  # BM25 keys: value descending
    for keys, is_desc in ((asc, False), (desc, True)):
        keys = np.concatenate([keys, np.array([0xFFFFFFFFFFFFFFFF], np.uint64)])
        v, o = bench._key_values(keys, is_desc)
        assert np.array_equal(v[:-1], vals) and np.array_equal(o[:-1], ords.astype(np.int64)) and o[-1] == -1
