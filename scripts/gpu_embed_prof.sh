#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests/test_gpu_embed.py -x -q 2>&1 | tail -5 | cut -c1-300
python scripts/embed_probe.py bge-base 256 32
python scripts/embed_probe.py bge-base 1 32
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -s 300 -c 90 --csv --log-file gpurun_out/launches_embed.csv python scripts/embed_probe.py bge-base 256 32 > /dev/null 2>&1
echo done
