#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
step() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.log; ( time timeout "$@" ) > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/summary.log; tail -n 8 gpurun_out/$name.log | cut -c1-300 | tee -a gpurun_out/summary.log; }
: > gpurun_out/summary.log
step pytest_gpu 1500 python -m pytest tests -m gpu -x -q
step bench_c3 900 python bench.py --no-cpu-baseline
echo done
