#!/bin/bash
# One gpurun call: smoke, GPU parity tests, bench, ncu launch list. Outputs under gpurun_out/.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
step() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.log; ( time timeout "$@" ) > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/summary.log; tail -n 12 gpurun_out/$name.log | tee -a gpurun_out/summary.log; }
: > gpurun_out/summary.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee -a gpurun_out/summary.log
step smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
step pytest_gpu 1500 python -m pytest tests -m gpu -x -q
step bench_tiny 600 python bench.py --workload tiny --steps 3 --warmup 3
step bench_c3 1200 python bench.py
step ncu_list 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline
step sanitizer 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()"
echo done
