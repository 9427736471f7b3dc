#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
step() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.log; ( time timeout "$@" ) > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/summary.log; tail -n 14 gpurun_out/$name.log | cut -c1-300 | tee -a gpurun_out/summary.log; }
: > gpurun_out/summary.log
step pytest_gpu 1500 python -m pytest tests -m gpu -x -q
step bench_c3 900 python bench.py
step ncu_list_c3 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
echo done
