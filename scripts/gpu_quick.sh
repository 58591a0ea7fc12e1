#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_train_kernels_gpu.py -q -m gpu 2>&1 | tail -25
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -q -m gpu -x 2>&1 | tail -5
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_q.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e'], d['roofline']['kernel_time_breakdown_ms'], d['roofline']['frac'])
PY
tail -3 gpurun_out/bench_q.err
