#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests/test_gpu_embed.py -x -q 2>&1 | tail -2
for b in 256 32 1; do python scripts/embed_probe.py bge-base $b 32; done
KRAG_EMBED_GRAPH=0 python scripts/embed_probe.py bge-base 1 32
timeout 600 python bench.py --no-cpu-baseline | python -c "
import sys,json
j=json.loads(sys.stdin.read()); r=j['roofline']
print('value',round(j['value'],1),'e2e',round(j['e2e']['value'],1),'ms/step',round(j['ms_per_step'],2),'dense',round(r['dense_stage_ms'],2),'bm25',round(r['bm25_stage_ms'],2),'embed',round(j['embed']['batch_ms'],2),'embed_b1',round(j['embed']['batch1_ms'],3),'b1',j['batch1'])"
