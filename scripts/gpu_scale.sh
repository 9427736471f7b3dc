#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-4}
: > gpurun_out/summary.log
nvidia-smi -L | head -8 | tee -a gpurun_out/summary.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_gpu_check.py 2>&1 | tail -2 | tee -a gpurun_out/summary.log
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then timeout 900 python bench.py --gpus 1 --no-cpu-baseline --no-optin > gpurun_out/scale_$n.log 2>&1
    else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --no-cpu-baseline --no-optin > gpurun_out/scale_$n.log 2>&1; fi
    echo "n=$n exit=$?" | tee -a gpurun_out/summary.log; grep '^{' gpurun_out/scale_$n.log | cut -c1-160 | tee -a gpurun_out/summary.log
  fi
done
