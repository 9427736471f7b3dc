#!/bin/bash
# small-M split-K GEMM path + programmatic dependent launch: tests, latency with and without, bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_embed.py -x -q 2>&1 | tail -15 | cut -c1-300
for b in 1 8 32; do for s in 16 32; do
  python scripts/embed_probe.py bge-base $b $s
  KRAG_PDL=0 python scripts/embed_probe.py bge-base $b $s | sed 's/^/   [no PDL] /'
done; done
KRAG_GEMM_SPLITK=0 python scripts/embed_probe.py bge-base 1 16 | sed 's/^/   [no split-K] /'
python scripts/embed_probe.py bge-base 256 32
timeout 600 python bench.py --no-cpu-baseline | python -c "
import sys,json
j=json.loads(sys.stdin.read()); r=j['roofline']
print('value',round(j['value'],1),'e2e',round(j['e2e']['value'],1),'ms/step',round(j['ms_per_step'],2),'dense',round(r['dense_stage_ms'],2),'bm25',round(r['bm25_stage_ms'],2),'embed',round(j['embed']['batch_ms'],2),'embed_b1',round(j['embed']['batch1_ms'],3),'b1',j['batch1'])"
