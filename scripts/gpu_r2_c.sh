#!/bin/bash
# round 2, call C (1 GPU): K3 v2b (cp.async staging + accumulator sweep), K5 precision numbers, ncu of the K5 pair GEMM
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/r2c
timeout 900 python -m pytest tests/test_gpu_embed.py -q -s 2>&1 | grep -E "max|passed|failed|Error|error|assert" | tail -60 | tee gpurun_out/r2c/pytest_embed.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_vector_store.py tests/test_service.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2c/pytest_parity.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-optin > gpurun_out/r2c/bench.json 2> gpurun_out/r2c/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2c/bench.json') if l.startswith('{')][-1])
print('value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e']['value'], 'embed', j['embed']['batch_ms'], j['embed']['batch1_ms'], 'k3', j['roofline_k3']['stage_ms'], j['roofline_k3']['frac'], 'dense', j['roofline']['dense_stage_ms'], 'check', j['check']['fused_ids_equal'], j['check']['recall_at_10'], 'b1', j['batch1'])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_warp_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/r2c/k3_main python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-optin --embedding none > gpurun_out/r2c/ncu_k3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm2_f16s_kernel --launch-skip 50 --launch-count 4 -o gpurun_out/r2c/k5_gemm2 python scripts/embed_probe.py bge-base 256 32 > gpurun_out/r2c/ncu_k5.log 2>&1
ls -la gpurun_out/r2c/
