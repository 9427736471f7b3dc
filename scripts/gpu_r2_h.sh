#!/bin/bash
# round 2, call H (1 GPU): LN in registers, K2 sampled minima, K3 dense staging + auto kernel choice, postings built in term ranges, HTTP fast path
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/r2h
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/r2h/pytest.log
timeout 300 python -m pytest tests/test_gpu_embed.py -q -s -k "bert_forward or gemm" 2>&1 | grep -E "^(gemm|linear|small|short|bge|base|large)[^ ]*:? " | tee gpurun_out/r2h/k5_precision.log | tail -8
for bs in "256 32" "32 32" "1 32"; do timeout 120 python scripts/embed_probe.py bge-base $bs; done 2>&1 | tee gpurun_out/r2h/embed_probe.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h/bench.json 2> gpurun_out/r2h/bench.err; rc=$?; echo "bench rc=$rc"
if [ $rc -ne 0 ]; then tail -5 gpurun_out/r2h/bench.err; echo "retry with the first-generation K3"; KRAG_BM25_KERNEL=legacy timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h/bench.json 2> gpurun_out/r2h/bench_legacy.err; echo "bench(legacy K3) rc=$?"; fi
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2h/bench.json') if l.startswith('{')][-1])
print('value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e']['value'], 'embed', j['embed']['batch_ms'], j['embed']['batch1_ms'], 'k3', j['roofline_k3']['stage_ms'], j['roofline_k3']['frac'], 'dense', j['roofline']['dense_stage_ms'], j['roofline']['kernel_ms'], j['roofline']['frac'], 'check', j['check']['fused_ids_equal'], j['check']['recall_at_10'], 'b1', j['batch1']['value'], 'fallbacks', j['config']['tc_certificate_fallback_queries'])
PY
nproc
timeout 600 python scripts/http_load.py --docs 10000000 --seconds 6 --clients 20 --concurrency 64 --http-workers 0,4,8 2> gpurun_out/r2h/http_load.err | grep '^{' | tee gpurun_out/r2h/http_load_n1.json | cut -c1-700
tail -3 gpurun_out/r2h/http_load.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2h/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-optin > gpurun_out/r2h/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2h/launches.csv 2>/dev/null | grep -v "synth\|df_hist\|cub::\|row_norms\|score_postings\|tile_\|pack_sort\|expand_entry\|at::" | head -28
