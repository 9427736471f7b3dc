#!/bin/bash
# round 2, call F (8 GPUs): the remaining BASELINE configs with the final kernels + multi-rank correctness + HTTP load at N=8
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/r2f
export KRAG_BENCH_TRACE=1
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "${@:2}"; }
line() { grep '^{' "$1" | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read())
r = j.get('roofline', {}); c = j.get('check') or {}
print('$2', 'value', round(j['value'], 1), j['unit'], 'ms/step', round(j.get('ms_per_step', 0), 3), 'e2e', round((j.get('e2e') or {}).get('value', 0), 1), 'b1', round((j.get('batch1') or {}).get('value', 0), 1),
      'dense_ms', r.get('dense_stage_ms'), 'kernel_ms', r.get('kernel_ms'), 'frac', r.get('frac'), 'bm25_ms', r.get('bm25_stage_ms'), 'embed_ms', (j.get('embed') or {}).get('batch_ms'),
      'check', c.get('fused_ids_equal'), c.get('dense_lists_equal_exact_scan'), 'recall', c.get('recall_at_10'))"; }
nvidia-smi -L | wc -l
timeout 400 bash -c "$(declare -f tr); tr 8 bench.py --gpus 8 --steps 20 --warmup 5" > gpurun_out/r2f/c3_n8.out 2> gpurun_out/r2f/c3_n8.err; echo "c3 n8 rc=$?"; line gpurun_out/r2f/c3_n8.out c3_n8
timeout 600 bash -c "$(declare -f tr); tr 8 bench.py --gpus 8 --workload headline --steps 5 --warmup 3 --no-optin" > gpurun_out/r2f/headline_n8.out 2> gpurun_out/r2f/headline_n8.err; echo "headline n8 rc=$?"; line gpurun_out/r2f/headline_n8.out headline_n8
timeout 600 bash -c "$(declare -f tr); tr 8 bench.py --gpus 8 --workload c4 --steps 5 --warmup 3 --no-optin" > gpurun_out/r2f/c4_n8.out 2> gpurun_out/r2f/c4_n8.err; echo "c4 n8 rc=$?"; line gpurun_out/r2f/c4_n8.out c4_n8
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -k "4" 2>&1 | tail -4 | tee gpurun_out/r2f/pytest_sharded.log
timeout 400 bash -c "$(declare -f tr); tr 8 bench_index.py --chunks 131072 --seq 256" 2> gpurun_out/r2f/index_n8.err | tail -1 | tee gpurun_out/r2f/index_n8.json | cut -c1-600
timeout 400 python scripts/http_load.py --gpus 8 --docs 10000000 --seconds 6 --clients 8 --concurrency 64 2> gpurun_out/r2f/http_n8.err | tail -1 | tee gpurun_out/r2f/http_load_n8.json | cut -c1-700
timeout 600 bash -c "$(declare -f tr); tr 4 bench.py --gpus 4 --workload headline --steps 5 --warmup 3 --no-optin" > gpurun_out/r2f/headline_n4.out 2> gpurun_out/r2f/headline_n4.err; echo "headline n4 rc=$?"; line gpurun_out/r2f/headline_n4.out headline_n4
tail -3 gpurun_out/r2f/*.err | cut -c1-300 | tail -40
ls gpurun_out/r2f/
