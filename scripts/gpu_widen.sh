#!/bin/bash
# chat route, filter pushdown, peer exchange (1 rank) and the whole GPU suite
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | cut -c1-300
