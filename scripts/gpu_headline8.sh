#!/bin/bash
# BASELINE.json headline corpus: 100M x 768 fp32 (+ postings) sharded over 8 B200s; plus the c3 scaling point at N=8
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
: > gpurun_out/summary.log
nvidia-smi -L | wc -l | tee -a gpurun_out/summary.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --no-cpu-baseline --no-optin > gpurun_out/scale_8.log 2>&1
echo "c3 n=8 exit=$?" | tee -a gpurun_out/summary.log; grep '^{' gpurun_out/scale_8.log | cut -c1-200 | tee -a gpurun_out/summary.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --workload headline --no-cpu-baseline --no-optin --steps 5 > gpurun_out/headline_8.log 2>&1
echo "headline n=8 exit=$?" | tee -a gpurun_out/summary.log; grep '^{' gpurun_out/headline_8.log | cut -c1-200 | tee -a gpurun_out/summary.log; tail -3 gpurun_out/headline_8.log | cut -c1-300 | tee -a gpurun_out/summary.log
