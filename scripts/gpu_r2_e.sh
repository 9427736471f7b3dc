#!/bin/bash
# round 2, call E (1 GPU): whole GPU suite, bench (both arms), K3 profile, HTTP load test
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/r2e
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2e/pytest.log
timeout 300 python -m pytest tests/test_gpu_embed.py -q -s -k "bert_forward or gemm" 2>&1 | grep -E "^(gemm|linear|small|short|bge|base|large)[^ ]*:? " | tee gpurun_out/r2e/k5_precision.log | tail -8
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e/bench.json 2> gpurun_out/r2e/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2e/bench.json') if l.startswith('{')][-1])
print('value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e']['value'], 'embed', j['embed']['batch_ms'], j['embed']['batch1_ms'], 'k3', j['roofline_k3']['stage_ms'], j['roofline_k3']['frac'], 'dense', j['roofline']['dense_stage_ms'], j['roofline']['kernel_ms'], j['roofline']['frac'], 'check', j['check']['fused_ids_equal'], j['check']['recall_at_10'], 'b1', j['batch1']['value'], 'cpu', j.get('cpu_baseline',{}).get('value'), j.get('cpu_baseline',{}).get('measured'))
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2e/bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/r2e/bench_reference.json
timeout 600 python scripts/http_load.py --docs 10000000 --seconds 6 --clients 8 --concurrency 64 2> gpurun_out/r2e/http_load.err | tail -1 | tee gpurun_out/r2e/http_load_n1.json | cut -c1-900
tail -3 gpurun_out/r2e/http_load.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_warp_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/r2e/k3_main python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-optin --embedding none > gpurun_out/r2e/ncu_k3.log 2>&1
timeout 300 python bench_index.py --chunks 8192 --seq 256 2>gpurun_out/r2e/bench_index.err | tail -1 | tee gpurun_out/r2e/bench_index_n1.json | cut -c1-700
ls -la gpurun_out/r2e/
