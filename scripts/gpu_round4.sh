#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
step() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.log; ( time timeout "$@" ) > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/summary.log; tail -n 8 gpurun_out/$name.log | cut -c1-300 | tee -a gpurun_out/summary.log; }
: > gpurun_out/summary.log
step bench_c3 900 python bench.py
step bench_c3_noembed 900 python bench.py --embedding none --no-cpu-baseline
step bench_ref 600 python bench.py --impl reference --steps 3 --warmup 1
step pytest_gpu 1500 python -m pytest tests -m gpu -x -q
step ncu_list_c3 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
echo done
