#!/bin/bash
# round 2, 8 GPUs: HTTP load through the sharded service (single-threaded CPU-side torch in the engine processes)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2n8c
mkdir -p $O
timeout 240 python scripts/http_load.py --gpus 8 --docs 10000000 --seconds 5 --clients 24 --concurrency 64 --http-workers 0,8 2> $O/http_n8.err | grep '^{' | tee $O/http_load_n8.json | cut -c1-560
tail -3 $O/http_n8.err | cut -c1-300
