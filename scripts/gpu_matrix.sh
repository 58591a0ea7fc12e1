#!/bin/bash
mkdir -p gpurun_out
for flags in "" "--no-pdl" "--no-ln-fold" "--no-pdl --no-ln-fold"; do
for s in 1 8; do
  timeout -s KILL 300 python bench.py --steps 12 --warmup 12 --slots $s --no-train --no-cpu-baseline $flags > gpurun_out/bm.json 2> gpurun_out/bm.err
  python -c "import json;d=json.load(open('gpurun_out/bm.json'));print('$flags', $s, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))"
done; done
