import sys, numpy as np
sys.path.insert(0, ".")
import bench
from kaito_b200 import _native
ctx = _native.Context(0)
name = sys.argv[1] if len(sys.argv) > 1 else "bge-base"
B, S = int(sys.argv[2]) if len(sys.argv) > 2 else 256, int(sys.argv[3]) if len(sys.argv) > 3 else 32
cfg = bench.BGE[name]
e = _native.Embedder(ctx, cfg["num_hidden_layers"], cfg["hidden_size"], cfg["num_attention_heads"], cfg["intermediate_size"], cfg["vocab_size"])
e.load_state_dict(bench.random_bert_state(cfg))
toks = [np.random.default_rng(b).integers(1000, 30000, S) for b in range(B)]
import time
for _ in range(3): out = e.embed(toks)
t = time.perf_counter()
for _ in range(5): out = e.embed(toks)
dt = (time.perf_counter() - t) / 5
print(f"{name} B={B} S={S}: {dt*1e3:.3f} ms/batch, {bench.bert_flops(cfg, S)*B/dt/1e12:.1f} TFLOP/s", out[0, :3])
