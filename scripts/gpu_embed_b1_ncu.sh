#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
KRAG_EMBED_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/embed_b1_launches.csv python scripts/embed_probe.py bge-base 1 16 > gpurun_out/embed_b1_ncu.log 2>&1
tail -2 gpurun_out/embed_b1_ncu.log
KRAG_EMBED_GRAPH=0 python scripts/embed_probe.py bge-base 1 16
python scripts/embed_probe.py bge-base 1 16
