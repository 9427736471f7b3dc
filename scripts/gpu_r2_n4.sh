#!/bin/bash
# round 2, 4 GPUs: the headline corpus (100M x 768 fp32 + 8.8e9 postings) on 4 GPUs -- 25M rows + 2.2e9 postings per GPU;
# fits since the postings are built one term range at a time
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/r2n4
export KRAG_BENCH_TRACE=1
# the build path changed (allocation retry, OOM diagnostics): parity of the postings first, on one GPU
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "bm25 or retrieve or persist" 2>&1 | tail -3 | tee gpurun_out/r2n4/pytest_bm25.log
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "${@:2}"; }
timeout 900 bash -c "$(declare -f tr); tr 4 bench.py --gpus 4 --workload headline --steps 5 --warmup 3 --no-optin --no-cpu-baseline" > gpurun_out/r2n4/headline_n4.out 2> gpurun_out/r2n4/headline_n4.err; echo "headline n4 rc=$?"
grep '^{' gpurun_out/r2n4/headline_n4.out | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read())
r = j.get('roofline', {}); c = j.get('check') or {}
print('headline_n4 value', round(j['value'], 1), j['unit'], 'ms/step', round(j.get('ms_per_step', 0), 3), 'e2e', round((j.get('e2e') or {}).get('value', 0), 1), 'b1', round((j.get('batch1') or {}).get('value', 0), 1),
      'dense_ms', r.get('dense_stage_ms'), 'kernel_ms', r.get('kernel_ms'), 'frac', r.get('frac'), 'bm25_ms', r.get('bm25_stage_ms'), 'embed_ms', (j.get('embed') or {}).get('batch_ms'),
      'check', c.get('fused_ids_equal'), c.get('dense_lists_equal_exact_scan'), 'recall', c.get('recall_at_10'))"
grep -E "free|out of device memory|Error" gpurun_out/r2n4/headline_n4.err | sort | uniq -c | sort -rn | head -12 | cut -c1-300
nvidia-smi --query-gpu=memory.used,memory.total --format=csv | head -5
