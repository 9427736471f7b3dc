#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -4 | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline --dense-mode 3 > gpurun_out/bench_bf16.log 2>&1; tail -1 gpurun_out/bench_bf16.log | cut -c1-250
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-150
