#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_tc_kernel -s 2 -c 2 -o gpurun_out/prof_tc python scripts/tc_probe.py 2000000 256 > gpurun_out/ncu_tc.log 2>&1
echo "ncu exit=$?"; tail -5 gpurun_out/ncu_tc.log
