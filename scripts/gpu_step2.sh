#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -q -m gpu -x > gpurun_out/t_inf.log 2>&1; echo "inference tests exit=$? $(tail -n 1 gpurun_out/t_inf.log)"
grep -E "^E  |Error|FAILED" gpurun_out/t_inf.log | head -20
timeout -s KILL 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -n 1 gpurun_out/smoke.log)"
for s in 1 4 8; do
  timeout -s KILL 300 python bench.py --steps 12 --warmup 12 --slots $s --no-train --no-cpu-baseline > gpurun_out/bench_s$s.json 2> gpurun_out/bench_s$s.err; echo "slots $s exit=$?"
  python -c "import json;d=json.load(open('gpurun_out/bench_s$s.json'));print($s, d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_time_breakdown_ms'])"; tail -3 gpurun_out/bench_s$s.err
done
