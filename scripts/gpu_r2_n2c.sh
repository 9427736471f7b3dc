#!/bin/bash
# round 2, 2 GPUs: sharded service with capacity-based staging buffers -- multi-rank tests + HTTP load
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2n2c
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_sharded.py -q 2>&1 | tail -4 | tee $O/pytest_sharded.log
timeout 300 python scripts/http_load.py --gpus 2 --docs 10000000 --seconds 6 --clients 20 --concurrency 64 --http-workers 0,8 2> $O/http_n2.err | grep '^{' | tee $O/http_load_n2.json | cut -c1-560
tail -3 $O/http_n2.err | cut -c1-300
