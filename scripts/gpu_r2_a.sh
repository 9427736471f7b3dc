#!/bin/bash
# round 2, call A (1 GPU): full GPU test suite with the new K3 + bench-shape K2 parity tests, bench line, launch list
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/r2a
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r2a/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/r2a/bench.json
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2a/bench.json') if l.startswith('{')][-1])
print('value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e']['value'])
print('k3',j.get('roofline_k3'))
print('dense',j['roofline']['dense_stage_ms'],j['roofline']['kernel_ms'],'embed',j['embed']['batch_ms'],'check',j['check'])
PY
KRAG_BM25_LEGACY=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-optin 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('LEGACY bm25 stage', j['roofline']['bm25_stage_ms'], 'value', j['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-optin > gpurun_out/r2a/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2a/launches.csv 2>/dev/null | head -40
