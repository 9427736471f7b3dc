#!/bin/bash
# round 2, call B (1 GPU): K5 split-fp16 GEMMs (tests + timing), ncu full of the K3 main pass
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/r2b
timeout 900 python -m pytest tests/test_gpu_embed.py -x -q -s 2>&1 | grep -E "max|passed|failed|Error|error|assert" | tail -60 | tee gpurun_out/r2b/pytest_embed.log
for bs in "256 32" "32 32" "1 32" "64 256"; do timeout 120 python scripts/embed_probe.py bge-base $bs; done 2>&1 | tee gpurun_out/r2b/embed_probe.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-optin > gpurun_out/r2b/bench.json 2> gpurun_out/r2b/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2b/bench.json') if l.startswith('{')][-1])
print('value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e']['value'], 'embed', j['embed']['batch_ms'], j['embed']['batch1_ms'], 'k3', j['roofline_k3']['stage_ms'], 'dense', j['roofline']['dense_stage_ms'], 'check', j['check']['fused_ids_equal'], j['check']['recall_at_10'])
PY
# ncu full: the MAIN pass of K3 is every second bm25_warp_kernel launch (launch-skip 3 = 2nd step's main pass)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_warp_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/r2b/k3_main python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-optin --embedding none > gpurun_out/r2b/ncu_k3.log 2>&1
ls -la gpurun_out/r2b/
