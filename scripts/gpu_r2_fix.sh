#!/bin/bash
# round 2 (2 GPUs): bench.py after the ShardedRetriever buffer change -- N = 1 and N = 2 lines with the in-bench check
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2fix
mkdir -p $O
show() { python - "$1" "$2" <<'PY'
import json, sys
j = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[2], 'value', round(j['value'], 1), 'ms', round(j['ms_per_step'], 3), 'e2e', round(j['e2e']['value'], 1), 'k3', j['roofline']['bm25_stage_ms'], 'dense', j['roofline']['dense_stage_ms'], j['roofline']['kernel_ms'],
      'embed', j['embed']['batch_ms'], 'check', j['check'], 'clocks', j['clocks']['sm_mhz'])
PY
}
CUDA_VISIBLE_DEVICES=0 timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-optin > $O/bench_n1.json 2> $O/bench_n1.err; echo "n1 rc=$?"; show $O/bench_n1.json n1
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-optin > $O/bench_n2.json 2> $O/bench_n2.err; echo "n2 rc=$?"; show $O/bench_n2.json n2
tail -3 $O/bench_n1.err $O/bench_n2.err | cut -c1-200
