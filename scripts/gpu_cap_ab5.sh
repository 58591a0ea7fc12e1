#!/bin/bash
# persistent-grid targets of the decode GEMMs (bench.py --dec-ctas wide,narrow)
for dc in 0,0 48,60 40,60 56,60 64,60 48,40 48,120 36,48; do
  python bench.py --no-cpu-baseline --no-train --steps 20 --warmup 5 --dec-ctas $dc 2> gpurun_out/e.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(f\"dec_ctas=$dc  dev {d['ms_per_step']:.3f} ms  e2e {d['e2e']['ms_per_step']:.3f} ms  bf16-host {d['e2e_bf16_host']['ms_per_step']:.3f}\")"
done
