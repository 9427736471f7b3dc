#!/bin/bash
# round 2, call G (2 GPUs): multi-rank correctness (sharded retriever + sharded service), bench N=2, HTTP load on 2 GPUs
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/r2g
export KRAG_BENCH_TRACE=1
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -12 | tee gpurun_out/r2g/pytest_sharded.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2g/c3_n2.out 2> gpurun_out/r2g/c3_n2.err; echo "c3 n2 rc=$?"
grep '^{' gpurun_out/r2g/c3_n2.out | python -c "
import sys, json
j = json.loads(sys.stdin.read()); r = j['roofline']; c = j['check']
print('value', j['value'], 'ms', j['ms_per_step'], 'e2e', j['e2e']['value'], 'dense', r['dense_stage_ms'], r['kernel_ms'], 'bm25', r['bm25_stage_ms'], 'embed', j['embed']['batch_ms'], 'check', c)"
tail -3 gpurun_out/r2g/c3_n2.err
timeout 400 python scripts/http_load.py --gpus 2 --docs 10000000 --seconds 6 --clients 8 --concurrency 64 2> gpurun_out/r2g/http_n2.err | tail -1 | tee gpurun_out/r2g/http_load_n2.json | cut -c1-1200
tail -5 gpurun_out/r2g/http_n2.err
timeout 400 python scripts/http_load.py --gpus 1 --docs 10000000 --seconds 6 --clients 8 --concurrency 64 2> gpurun_out/r2g/http_n1.err | tail -1 | tee gpurun_out/r2g/http_load_n1.json | cut -c1-1200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench_index.py --chunks 16384 --seq 256 2> gpurun_out/r2g/index_n2.err | tail -1 | tee gpurun_out/r2g/index_n2.json | cut -c1-900
tail -3 gpurun_out/r2g/index_n2.err
