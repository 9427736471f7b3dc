#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "bm25 or retrieve or full_size" 2>&1 | tail -3 | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline --no-optin | python -c "
import sys,json
j=json.loads(sys.stdin.read()); r=j['roofline']
print('value',round(j['value'],1),'e2e',round(j['e2e']['value'],1),'ms/step',round(j['ms_per_step'],2),'dense',round(r['dense_stage_ms'],2),'bm25',round(r['bm25_stage_ms'],3),'embed',round(j['embed']['batch_ms'],2),'embed_b1',round(j['embed']['batch1_ms'],3),'b1',round(j['batch1']['value'],1))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bm25_tile_kernel -s 1 -c 1 -f -o gpurun_out/k3_now python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-optin > gpurun_out/k3_now.log 2>&1
