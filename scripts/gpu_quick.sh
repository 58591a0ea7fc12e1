#!/bin/bash
# quick loop: GEMM parity tests + decode kernel timings (+ optional bench)
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout -s KILL 400 python -m pytest tests/test_kernels_gpu.py tests/test_train_kernels_gpu.py -q -x -k "linear or gemm or wgrad or topk or attn or attention or box" 2>&1 | tail -3
timeout -s KILL 300 python scripts/dec_kernels.py --hints 20003256,3256 > gpurun_out/dec_kernels3.log 2>&1; grep -E "gemm|topk|cross|layernorm|self" gpurun_out/dec_kernels3.log
if [ "$1" = "bench" ]; then
  python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
  python - <<'PY'
import json
d=json.load(open("gpurun_out/q_bench.json")); r=d["roofline"]; t=d.get("train") or {}
print(f"dev {d['ms_per_step']:.3f} ms  e2e {d['e2e']['ms_per_step']:.3f} ms  gemm alone {r['achieved']:.0f} in-flight {r['achieved_in_flight']:.0f}  train {t.get('ms_per_step')}")
PY
fi
