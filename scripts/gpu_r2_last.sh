#!/bin/bash
# round 2, last call (1 GPU): the whole GPU suite on the final tree, the bench line, HTTP load with the final host
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2last
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/pytest.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-optin > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.loads([l for l in open('gpurun_out/r2last/bench.json') if l.startswith('{')][-1])
print('value', round(j['value'], 1), 'ms', round(j['ms_per_step'], 3), 'e2e', round(j['e2e']['value'], 1), 'k3', j['roofline_k3']['stage_ms'], 'dense', j['roofline']['dense_stage_ms'], j['roofline']['kernel_ms'], 'embed', j['embed']['batch_ms'], 'check', j['check']['fused_ids_equal'], 'clocks', j['clocks']['sm_mhz'])
PY
timeout 200 python scripts/http_load.py --docs 10000000 --seconds 5 --clients 24 --concurrency 64 --http-workers 0,8 2> $O/http.err | grep '^{' | tee $O/http_load_n1.json | cut -c1-420
tail -2 $O/http.err | cut -c1-200
