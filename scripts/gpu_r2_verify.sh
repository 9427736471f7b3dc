#!/bin/bash
# round 2, verification of the driver's commands on the final tree (1 GPU): smoke(), the default bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2v
mkdir -p $O
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.loads([l for l in open('gpurun_out/r2v/bench.json') if l.startswith('{')][-1])
print('value', round(j['value'], 1), 'ms', round(j['ms_per_step'], 3), 'e2e', round(j['e2e']['value'], 1), 'steps', j['steps'], j['warmup'], 'check', j['check']['fused_ids_equal'], j['recall_at_10'],
      'optin', (j.get('optin_bf16_shadow') or {}).get('value'), 'cpu', (j.get('cpu_baseline') or {}).get('value'), 'launches', j['gpu_launches'], 'clocks', j['clocks'])
PY
tail -2 $O/bench.err | cut -c1-200
