#!/bin/bash
# peer-memory exchange: parity vs the NCCL path + exchange cost, then the N-GPU bench with both exchanges
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
step() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.log; ( time timeout "$@" ) > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/summary.log; tail -n 8 gpurun_out/$name.log | cut -c1-600 | tee -a gpurun_out/summary.log; }
: > gpurun_out/summary.log
nvidia-smi -L | tee -a gpurun_out/summary.log
nvidia-smi topo -m 2>&1 | head -12 | tee -a gpurun_out/summary.log
step sharded_check_$N 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_gpu_check.py
step bench_p2p_$N 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --no-cpu-baseline --steps 10 --warmup 3
KRAG_P2P=0 step bench_nccl_$N 600 env KRAG_P2P=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --no-cpu-baseline --steps 10 --warmup 3
echo done
