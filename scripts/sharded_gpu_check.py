"""torchrun --nproc-per-node N scripts/sharded_gpu_check.py : N-GPU document-sharded /retrieve
(CUDA stages + one NCCL all-gather) must equal the single-shard oracle over the whole corpus."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from kaito_b200 import _native
from kaito_b200.sharded import NativeStages, ShardedRetriever, shard_range
from oracle import oracle as o

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n, d, vocab, k, B = 50_000, 96, 4000, 10, 12
x = o.synth_dense(n, d, 1)
off, ids, tf, dl = o.synth_sparse(n, vocab, 2)
q = o.synth_queries(x, B, 3)
qs = o.synth_query_terms(vocab, B, 4, rank_offset=30)
lo, hi = shard_range(n, world, rank)
ctx = _native.Context(device_id=lr, rank=rank, world_size=world)
ix = ctx.create_index("shard", d)
ix.add(np.arange(lo, hi, dtype=np.uint64), x[lo:hi], off[lo:hi + 1] - off[lo], ids[off[lo]:off[hi]], tf[off[lo]:off[hi]], dl[lo:hi])
sr = ShardedRetriever(NativeStages(ctx, ix), dev, ix.stats().dim_padded)
n_docs, total, base = sr.commit(vocab, hi - lo)
assert (n_docs, base) == (n, lo)
got = sr.retrieve(q, qs, k)
got2 = sr.retrieve(q, None, k)
post = o.bm25_build(off, ids, tf, dl, vocab)
P = o.pool_size(k)
for b in range(B):
    dd, do = o.dense_topk(x, q[b:b + 1], P)
    bs, bo = o.bm25_query(post, qs[b], P)
    fin, de, sp, rk, od = o.fuse(dd[0], do[0], bs, bo, k)
    c = int(got["count"][b])
    assert c == len(od), (c, len(od))
    assert np.array_equal(got["ordinal"][b, :c], od), (rank, b, got["ordinal"][b, :c], od)
    assert np.array_equal(got["final"][b, :c], fin)
    assert np.array_equal(got2["ordinal"][b], do[0, :k])
# the peer-memory exchange (default) and the NCCL all-gather path must agree, over repeated exchanges (mailbox parity
# slots) and changing batch sizes
sn = ShardedRetriever(NativeStages(ctx, ix), dev, ix.stats().dim_padded)
sn.hybrid, sn._p2p_state = True, False
assert world == 1 or sr._p2p_state is True, "peer-memory exchange did not come up"
for it in range(6):
    b = [B, 1, 5, B, 3, 7][it]
    a1 = sr.retrieve(q[:b], qs[:b] if it % 2 == 0 else None, k)
    a2 = sn.retrieve(q[:b], qs[:b] if it % 2 == 0 else None, k)
    for key in ("ordinal", "final", "count"):
        assert np.array_equal(a1[key], a2[key]), (rank, it, key)
# exchange cost: retrieve_dev with both paths, batch 1 and 256 (device time, max over ranks)
import time
def timed(r, b, hybrid, iters=30):
    qd = torch.from_numpy(np.pad(o.synth_queries(x, b, 9), ((0, 0), (0, r.dpad - d)))).to(dev)
    tq = o.synth_query_terms(vocab, b, 10, rank_offset=30)
    flat = np.concatenate(tq).astype(np.int32); toff_h = np.zeros(b + 1, np.int32); toff_h[1:] = np.cumsum([len(t) for t in tq])
    terms, toff = torch.from_numpy(flat).to(dev), torch.from_numpy(toff_h).to(dev)
    for _ in range(5):
        r.retrieve_dev(qd, terms if hybrid else None, toff if hybrid else None, k, toff_host=toff_h)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        r.retrieve_dev(qd, terms if hybrid else None, toff if hybrid else None, k, toff_host=toff_h)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
if world > 1:
    for b in (1, 256):
        tp, tn = timed(sr, b, True), timed(sn, b, True)
        if rank == 0:
            print(f"exchange batch {b}: peer-memory {tp*1e3:.1f} us/step, NCCL all-gather + merges {tn*1e3:.1f} us/step")
dist.barrier()
if rank == 0:
    print(f"sharded gpu check ok: world={world}, ids/finals bit-exact vs single-shard oracle; p2p == nccl path")
ix.drop(); ctx.close()
dist.destroy_process_group()
