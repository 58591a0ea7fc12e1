#!/bin/bash
for i in 1 2; do python bench.py --no-train --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/e.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'], d['e2e_bf16_host']['ms_per_step'])"; done
python bench.py --config scst --no-cpu-baseline --steps 20 --warmup 5 2>>gpurun_out/e.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'], d['e2e_bf16_host']['ms_per_step'])"
tail -3 gpurun_out/e.err
