#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
: > gpurun_out/summary.log
echo "=== tc tests (2cta)" | tee -a gpurun_out/summary.log
KRAG_TC_2CTA=1 timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -15 | cut -c1-300 | tee -a gpurun_out/summary.log
echo "=== bench 2cta" | tee -a gpurun_out/summary.log
KRAG_TC_2CTA=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_2cta.log 2>&1; tail -2 gpurun_out/bench_2cta.log | cut -c1-300 | tee -a gpurun_out/summary.log
echo "=== ncu 2cta" | tee -a gpurun_out/summary.log
KRAG_TC_2CTA=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dense_tc2_kernel" -s 1 -c 1 -o gpurun_out/prof_2cta python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_2cta.log 2>&1; tail -3 gpurun_out/ncu_2cta.log | cut -c1-200 | tee -a gpurun_out/summary.log
