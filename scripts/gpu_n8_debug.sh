#!/bin/bash
# round 2: reproduce the driver's N=8 launch of bench.py exactly (SCALE_r01: rank 0 SIGABRT) with stderr, Python
# stacks (faulthandler), C++ stack traces and phase markers kept; on failure bisect with the two obvious knobs.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/n8
export KRAG_BENCH_TRACE=1 TORCH_SHOW_CPP_STACKTRACES=1 NCCL_DEBUG=WARN
run() {   # name, extra env (as VAR=VAL words), extra args
  local name=$1; shift
  local envs=$1; shift
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 295$((RANDOM % 90 + 10)) \
      bench.py --gpus 8 --steps 20 --warmup 5 "$@" > gpurun_out/n8/$name.out 2> gpurun_out/n8/$name.err
  local rc=$?
  echo "== $name rc=$rc"; grep '^{' gpurun_out/n8/$name.out | cut -c1-300
  grep -v "^\[bench rank [1-7]" gpurun_out/n8/$name.err | grep -i -B2 -A25 "abort\|terminate\|error\|p2p_merge\|Fatal" | head -80
  grep "^\[bench rank 0" gpurun_out/n8/$name.err | tail -3
  return $rc
}
nvidia-smi -L | wc -l
if run exact "A=1"; then
  run exact2 "A=1"
else
  run nooptin "A=1" --no-optin
  run nccl "KRAG_P2P=0"
fi
dmesg 2>/dev/null | tail -5
