#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q  2>&1 | tail -12 | cut -c1-400
timeout 600 python bench.py --no-cpu-baseline --no-optin | python -c "
import sys,json
j=json.loads(sys.stdin.read()); r=j['roofline']
print('value',round(j['value'],1),'e2e',round(j['e2e']['value'],1),'ms/step',round(j['ms_per_step'],2),'dense',round(r['dense_stage_ms'],2),'kern',round(r['kernel_ms'],3),'frac',round(r['frac'],3),'bm25',round(r['bm25_stage_ms'],3),'embed',round(j['embed']['batch_ms'],2),'fb',j['config']['tc_certificate_fallback_queries'], j['config']['dense_kernel'][:40], j['clocks'])"
