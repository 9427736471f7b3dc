#!/bin/bash
# round 2, final evidence (1 GPU): bench both arms, c2 line, /index bench, ncu --set full of the top kernels, launch list
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2z
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee $O/pytest.log
timeout 300 python -m pytest tests/test_gpu_embed.py -q -s -k "bert_forward or gemm" 2>&1 | grep -E "^(gemm|linear|small|short|bge|base|large)[^ ]*:? " | tee $O/k5_precision.log | tail -8
for bs in "256 32" "32 32" "1 32" "128 256"; do timeout 120 python scripts/embed_probe.py bge-base $bs; done 2>&1 | tee $O/embed_probe.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2>/dev/null; echo "reference rc=$?"; cut -c1-260 $O/bench_reference.json
timeout 600 python bench.py --workload c2 --steps 20 --warmup 5 --no-optin > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
for f in ('bench', 'bench_c2'):
    try:
        j = json.loads([l for l in open(f'gpurun_out/r2z/{f}.json') if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'no line', e); continue
    r, k3, c = j.get('roofline', {}), j.get('roofline_k3') or {}, j.get('check') or {}
    print(f, 'value', round(j['value'], 1), 'ms', round(j['ms_per_step'], 3), 'e2e', round(j['e2e']['value'], 1), 'embed', (j.get('embed') or {}).get('batch_ms'),
          'k3', k3.get('stage_ms'), k3.get('frac'), 'dense', r.get('dense_stage_ms'), r.get('kernel_ms'), r.get('frac'), 'check', c.get('fused_ids_equal'), c.get('recall_at_10'),
          'b1', (j.get('batch1') or {}).get('value'), 'cpu', (j.get('cpu_baseline') or {}).get('value'), 'clocks', j.get('clocks'))
PY
timeout 300 python bench_index.py --chunks 8192 --seq 256 2> $O/bench_index.err | tail -1 | tee $O/bench_index_n1.json | cut -c1-700
NCU="ncu --set full --clock-control none --import-source on --launch-count 1"
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-optin"
timeout 600 $NCU -k regex:dense_tc2cvt_kernel --launch-skip 2 -o $O/k2_main $B > $O/ncu_k2.log 2>&1
timeout 600 $NCU -k regex:bm25_warp_kernel --launch-skip 3 -o $O/k3_main $B --embedding none > $O/ncu_k3.log 2>&1
timeout 600 $NCU -k regex:gemm2_f16s_kernel --launch-skip 100 -o $O/k5_gemm2 $B > $O/ncu_k5.log 2>&1
timeout 600 $NCU -k regex:attention_mma_kernel --launch-skip 30 -o $O/k5_attn $B > $O/ncu_k5a.log 2>&1
python scripts/summarize_ncu.py $O/ncu_full.txt $O/k2_main.ncu-rep $O/k3_main.ncu-rep $O/k5_gemm2.ncu-rep $O/k5_attn.ncu-rep 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches.csv $B > $O/ncu_bench.log 2>&1
python scripts/summarize_launches.py $O/launches.csv 2>/dev/null | grep -v "synth\|df_hist\|cub::\|row_norms\|score_postings\|tile_\|chunk_\|expand_entry\|at::" | head -30 | tee $O/launches_summary.txt
ls -la $O
