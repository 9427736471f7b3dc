#!/bin/bash
# round 2, 8 GPUs, final code: c3 bench line (the driver's command) + HTTP load through the sharded service with front-end workers
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2n8
mkdir -p $O
export KRAG_BENCH_TRACE=1
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "${@:2}"; }
timeout 300 bash -c "$(declare -f tr); tr 8 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline" > $O/c3_n8.out 2> $O/c3_n8.err; echo "c3 n8 rc=$?"
grep '^{' $O/c3_n8.out | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read()); r = j.get('roofline', {}); c = j.get('check') or {}
print('c3_n8 value', round(j['value'], 1), 'ms/step', round(j['ms_per_step'], 3), 'e2e', round(j['e2e']['value'], 1), 'b1', round(j['batch1']['value'], 1), 'dense_ms', r.get('dense_stage_ms'), 'kernel_ms', r.get('kernel_ms'),
      'bm25_ms', r.get('bm25_stage_ms'), 'embed_ms', (j.get('embed') or {}).get('batch_ms'), 'check', c.get('fused_ids_equal'), c.get('recall_at_10'), 'clocks', j['clocks']['sm_mhz'])"
for w in headline c4; do
  timeout 240 bash -c "$(declare -f tr); tr 8 bench.py --gpus 8 --workload $w --steps 5 --warmup 3 --no-optin --no-cpu-baseline" > $O/${w}_n8.out 2> $O/${w}_n8.err; echo "$w n8 rc=$?"
  grep '^{' $O/${w}_n8.out | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read()); r = j.get('roofline', {}); c = j.get('check') or {}
print('$w n8 value', round(j['value'], 1), 'ms/step', round(j['ms_per_step'], 3), 'e2e', round(j['e2e']['value'], 1), 'b1', round(j['batch1']['value'], 1), 'kernel_ms', r.get('kernel_ms'), 'frac', r.get('frac'),
      'bm25_ms', r.get('bm25_stage_ms'), 'embed_ms', (j.get('embed') or {}).get('batch_ms'), 'check', c.get('fused_ids_equal'), c.get('dense_lists_equal_exact_scan'), c.get('recall_at_10'), 'clocks', j['clocks']['sm_mhz'])"
done
timeout 300 python scripts/http_load.py --gpus 8 --docs 10000000 --seconds 6 --clients 24 --concurrency 64 --http-workers 0,8 2> $O/http_n8.err | grep '^{' | tee $O/http_load_n8.json | cut -c1-600
tail -3 $O/http_n8.err | cut -c1-300
tail -3 $O/c3_n8.err | cut -c1-300
