#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
step() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.log; ( time timeout "$@" ) > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/summary.log; tail -n 14 gpurun_out/$name.log | cut -c1-300 | tee -a gpurun_out/summary.log; }
: > gpurun_out/summary.log
step pytest_gpu 1500 python -m pytest tests -m gpu -x -q
step bench_c3 900 python bench.py
step ncu_tc 900 ncu --set full --clock-control none --import-source on -k regex:"dense_tc_kernel|bm25_tile_kernel" -s 3 -c 3 -o gpurun_out/prof_r3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline
echo done
