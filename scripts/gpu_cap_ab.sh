#!/bin/bash
for cap in ${CAPS:-0 128 112 96 0}; do
  SC_GEMM_MAX_CTAS=$cap python bench.py --no-cpu-baseline --no-train --steps 20 --warmup 5 2> gpurun_out/e.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(f\"cap=$cap  bench dev {d['ms_per_step']:.3f} ms  e2e {d['e2e']['ms_per_step']:.3f} ms\")"
done
