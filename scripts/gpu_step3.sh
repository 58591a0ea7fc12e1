#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu -x > gpurun_out/tests.log 2>&1; echo "tests exit=$? $(tail -n 1 gpurun_out/tests.log)"
grep -E "^E  |Error|FAILED" gpurun_out/tests.log | head -20
for flags in "" "--no-ln-fold"; do
for s in 1 8; do
  timeout -s KILL 300 python bench.py --steps 12 --warmup 12 --slots $s --no-train --no-cpu-baseline $flags > gpurun_out/bm.json 2> gpurun_out/bm.err
  python -c "import json;d=json.load(open('gpurun_out/bm.json'));print('$flags', $s, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))"
done; done
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --slots 8 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(d['ms_per_step'], d['e2e'], d['roofline']['frac']); print(d['train'])"
