#!/usr/bin/env python
"""bench_index.py -- the /index path (BASELINE.json configs[4]) on ONE GPU:
batch-embed chunks with K5 (bge-base shapes, random-init weights), append rows + term lists to the index,
commit (postings + tile index).  Reports chunks/s per stage and end to end, and the K5 tensor fraction.
Not the driver's bench (that is bench.py, /retrieve); numbers from this script go to profiles/ and DESIGN.md."""
import argparse
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from kaito_b200 import _native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chunks", type=int, default=16384)
ap.add_argument("--seq", type=int, default=256)
ap.add_argument("--model", default="bge-base")
ap.add_argument("--embed-batch", type=int, default=128, help="chunks per K5 call")
a = ap.parse_args()

cfg = bench.BGE[a.model]
ctx = _native.Context(0)
emb = _native.Embedder(ctx, cfg["num_hidden_layers"], cfg["hidden_size"], cfg["num_attention_heads"], cfg["intermediate_size"], cfg["vocab_size"])
emb.load_state_dict(bench.random_bert_state(cfg))
g = np.random.default_rng(0)
toks = [g.integers(1000, 30000, a.seq) for _ in range(a.embed_batch)]
emb.embed(toks)                                   # warm-up
t0 = time.perf_counter()
vecs = []
for b0 in range(0, a.chunks, a.embed_batch):
    vecs.append(emb.embed(toks[: min(a.embed_batch, a.chunks - b0)]))
t_embed = time.perf_counter() - t0
vecs = np.concatenate(vecs)
# sparse side: ~80 unique terms per chunk from a 2^18 vocabulary (host tokenisation is not timed here)
vocab = 1 << 18
offs = np.arange(0, (a.chunks + 1) * 80, 80, dtype=np.int64)
tids = g.integers(0, vocab, a.chunks * 80).astype(np.uint32)
tf = np.ones(a.chunks * 80, np.uint16)
dl = np.full(a.chunks, 96, np.uint32)
ix = ctx.create_index("idx", cfg["hidden_size"])
t0 = time.perf_counter()
ix.add(np.arange(a.chunks, dtype=np.uint64), vecs, offs, tids, tf, dl)
t_add = time.perf_counter() - t0
t0 = time.perf_counter()
ix.commit(vocab)
t_commit = time.perf_counter() - t0
fl = bench.bert_flops(cfg, a.seq) * a.chunks
peaks = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {}
tf32_peak = float(peaks.get("bf16_tflops", 1590.0)) / 2
print(json.dumps({
    "metric": "index_chunks_per_sec", "unit": "chunks/s", "n_gpus": 1, "data": "synthetic",
    "config": {"workload": f"/index: {a.chunks} chunks x {a.seq} tokens, {a.model} shapes (random-init), 80 terms/chunk, vocab 2^18",
               "embed_batch": a.embed_batch},
    "value": a.chunks / (t_embed + t_add + t_commit),
    "embed": {"chunks_per_s": a.chunks / t_embed, "tflops": fl / t_embed / 1e12, "tensor_frac_of_tf32_peak": fl / t_embed / 1e12 / tf32_peak,
              "flops_formula": "L*(24*S*d^2 + 4*S^2*d) per chunk (SURVEY.md 8d)"},
    "add_chunks_per_s": a.chunks / t_add, "commit_s": t_commit,
}))
ix.drop(); emb.destroy(); ctx.close()
