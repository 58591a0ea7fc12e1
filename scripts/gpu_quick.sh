#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_trainer_gpu.py -q -m gpu 2>&1 | tail -3
timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_q.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'])
print(d['train'])
PY
tail -5 gpurun_out/bench_q.err
