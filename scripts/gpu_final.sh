#!/bin/bash
# Round evidence: tests, bench (both arms), ncu --set full of the four dominant kernels, launch list.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
step() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.log; ( time timeout "$@" ) > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/summary.log; tail -n 6 gpurun_out/$name.log | cut -c1-300 | tee -a gpurun_out/summary.log; }
: > gpurun_out/summary.log
step smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
step pytest_gpu 1500 python -m pytest tests -m gpu -x -q
step bench_ref 600 python bench.py --impl reference
step bench_c3 900 python bench.py
step ncu_k2 600 ncu --set full --clock-control none --import-source on -k regex:dense_tc2cvt_kernel -s 1 -c 1 -f -o gpurun_out/final_k2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-optin
step ncu_k1 600 ncu --set full --clock-control none --import-source on -k regex:dense_scan_kernel -s 2 -c 1 -f -o gpurun_out/final_k1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-optin
step ncu_k3 600 ncu --set full --clock-control none --import-source on -k regex:bm25_tile_kernel -s 1 -c 1 -f -o gpurun_out/final_k3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-optin
step ncu_k5 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_tf32_kernel -s 50 -c 4 -f -o gpurun_out/final_k5 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-optin
step ncu_list 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-optin
echo done
