#!/bin/bash
# round 2, call I (1 GPU): K3 A/B on ONE box -- dense staging (shipped) vs the previous per-term rounds vs dense staging without the
# lane-parallel short-list copies; alternative libraries are prebuilt under kaito_b200/alt/ (same sources except bm25.cu)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2i
mkdir -p $O
cp kaito_b200/libkaito_rag.so /tmp/krag_main.so
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-optin"
show() { python - "$1" "$2" <<'PY'
import json, sys
j = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[2], 'value', round(j['value'], 1), 'ms', round(j['ms_per_step'], 3), 'k3', round(j['roofline_k3']['stage_ms'], 3), 'dense', round(j['roofline']['dense_stage_ms'], 3),
      round(j['roofline']['kernel_ms'], 3), 'embed', round(j['embed']['batch_ms'], 3), 'check', j['check']['fused_ids_equal'], 'clocks', j['clocks']['sm_mhz'], j['clocks']['reasons'])
PY
}
for v in main k3v2e k3v3s0 tcold main; do
  if [ $v = main ]; then cp /tmp/krag_main.so kaito_b200/libkaito_rag.so; else cp kaito_b200/alt/libkaito_rag_$v.so kaito_b200/libkaito_rag.so; fi
  timeout 600 $B > $O/bench_$v.json 2> $O/bench_$v.err; echo "bench $v rc=$?"
  show $O/bench_$v.json $v
done
for v in main k3v2e; do
  if [ $v = main ]; then cp /tmp/krag_main.so kaito_b200/libkaito_rag.so; else cp kaito_b200/alt/libkaito_rag_$v.so kaito_b200/libkaito_rag.so; fi
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:bm25_warp_kernel --launch-skip 3 --launch-count 1 -o $O/k3_$v python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-optin --embedding none > $O/ncu_k3_$v.log 2>&1
done
cp /tmp/krag_main.so kaito_b200/libkaito_rag.so
ls -la $O
