#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests/test_gpu_embed.py -x -q 2>&1 | tail -2
for b in 256 64 32 1; do python scripts/embed_probe.py bge-base $b 32; done
KRAG_GEMM_2CTA=0 python scripts/embed_probe.py bge-base 256 32
timeout 600 python bench_index.py --chunks 8192 --seq 256
