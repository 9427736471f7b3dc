#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
step() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.log; ( time timeout "$@" ) > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/summary.log; tail -n 6 gpurun_out/$name.log | cut -c1-400 | tee -a gpurun_out/summary.log; }
: > gpurun_out/summary.log
nvidia-smi -L | tee -a gpurun_out/summary.log
step sharded_check_$N 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_gpu_check.py
step bench_ref 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1
step bench_1 900 python bench.py --gpus 1 --no-cpu-baseline
step bench_$N 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --no-cpu-baseline
echo done
