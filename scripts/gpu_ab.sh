#!/bin/bash
# same-box A/B of two runtime knobs: BM25 group size, wide split-K tiles for long-K layers
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_embed.py -x -q -k "gemm or linear" 2>&1 | tail -2
for b in 32 16 64; do
  python scripts/embed_probe.py bge-base $b 32
  KRAG_SK_WIDE=0 python scripts/embed_probe.py bge-base $b 32 | sed 's/^/   [BN=64] /'
done
for g in 8 16 32; do
KRAG_BM25_GROUP=$g timeout 600 python bench.py --no-cpu-baseline --no-optin --steps 20 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); r=j['roofline']
print('group $g: value',round(j['value'],1),'ms/step',round(j['ms_per_step'],2),'dense',round(r['dense_stage_ms'],2),'bm25',round(r['bm25_stage_ms'],3),'embed',round(j['embed']['batch_ms'],2))"
done
KRAG_BM25_GROUP=16 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "bm25 or retrieve" 2>&1 | tail -2
