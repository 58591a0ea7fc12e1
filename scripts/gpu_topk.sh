#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -q -m gpu > gpurun_out/t_topk.log 2>&1; echo "tests exit=$? $(tail -n 1 gpurun_out/t_topk.log)"
grep -E "^E  |Error|FAILED" gpurun_out/t_topk.log | head -30
timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -1
for flags in "" "--no-fuse-topk"; do
  timeout -s KILL 300 python bench.py --steps 16 --warmup 8 --no-train --no-cpu-baseline $flags > gpurun_out/bm.json 2> gpurun_out/bm.err
  python -c "import json;d=json.load(open('gpurun_out/bm.json'));print('$flags', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), round(d['roofline']['achieved'],1), d['roofline']['dominant_shape'])" || tail -5 gpurun_out/bm.err
done
