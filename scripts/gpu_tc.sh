#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
step() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.log; ( time timeout "$@" ) > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/summary.log; tail -n 25 gpurun_out/$name.log | cut -c1-400 | tee -a gpurun_out/summary.log; }
: > gpurun_out/summary.log
step tc_raw 300 python -m pytest tests/test_gpu_tc.py -x -q -k raw
step tc_pipe 600 python -m pytest tests/test_gpu_tc.py -x -q -k "pipeline or certificate"
step bench_c3_tc 900 python bench.py --no-cpu-baseline
step pytest_gpu 1500 python -m pytest tests -m gpu -x -q
echo done
