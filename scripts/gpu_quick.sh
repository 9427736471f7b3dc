#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "raw or pipeline" 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; tail -1 gpurun_out/bench_quick.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpc__cycles_elapsed.max.per_second,dram__bytes_read.sum --clock-control none -k regex:"dense_tc_kernel" -s 3 -c 2 --csv --log-file gpurun_out/quick_tc.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
grep -v "^==" gpurun_out/quick_tc.csv | cut -d, -f5,13,15 | cut -c1-160
