#!/bin/bash
# LayerNorm folded into the consuming GEMMs (now with the engine's tile hints, the fast residual-stream producer epilogue and the
# fused generator kept) vs separate LayerNorm kernels
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -q -x 2>&1 | tail -2
for v in "" "--ln-fold" "" "--ln-fold"; do
  python bench.py --no-train --no-cpu-baseline --steps 20 --warmup 5 $v 2>gpurun_out/e.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']; print('fold=$v', 'dev', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), 'bf16-host', round(d['e2e_bf16_host']['ms_per_step'], 3), 'gemm', round(r['achieved']), round(r['achieved_in_flight']))"
done
tail -2 gpurun_out/e.err
