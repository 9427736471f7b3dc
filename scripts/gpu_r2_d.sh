#!/bin/bash
# round 2, call D (1 GPU): whole GPU suite (K3 v2c, K5 BN=128 double-buffered TMEM, sharded service at world 1), bench, profiles
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/r2d
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2d/pytest.log
timeout 300 python -m pytest tests/test_gpu_embed.py -q -s -k "bert_forward or gemm" 2>&1 | grep -E "^(gemm|linear|small|short|bge|base|large)[^ ]*:? " | tee gpurun_out/r2d/k5_precision.log | tail -12
for bs in "256 32" "32 32" "1 32" "64 256"; do timeout 120 python scripts/embed_probe.py bge-base $bs; done 2>&1 | tee gpurun_out/r2d/embed_probe.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d/bench.json 2> gpurun_out/r2d/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2d/bench.json') if l.startswith('{')][-1])
print('value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e']['value'], 'embed', j['embed']['batch_ms'], j['embed']['batch1_ms'], 'k3', j['roofline_k3']['stage_ms'], j['roofline_k3']['frac'], 'dense', j['roofline']['dense_stage_ms'], j['roofline']['kernel_ms'], j['roofline']['frac'], 'check', j['check']['fused_ids_equal'], j['check']['recall_at_10'], 'b1', j['batch1']['value'], 'cpu', j.get('cpu_baseline',{}).get('value'), j.get('cpu_baseline',{}).get('measured'))
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2d/bench_reference.json 2>/dev/null; cut -c1-400 gpurun_out/r2d/bench_reference.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_warp_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/r2d/k3_main python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-optin --embedding none > gpurun_out/r2d/ncu_k3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm2_f16s_kernel --launch-skip 48 --launch-count 4 -o gpurun_out/r2d/k5_gemm2 python scripts/embed_probe.py bge-base 256 32 > gpurun_out/r2d/ncu_k5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2d/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-optin > gpurun_out/r2d/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2d/launches.csv 2>/dev/null | grep -v "synth\|df_hist\|cub::\|row_norms\|score_postings\|tile_\|pack_sort\|expand_entry\|at::" | head -30
ls -la gpurun_out/r2d/
