#!/bin/bash
# end-to-end arm: queued batches per launch (the H2D copy of a launch's batches paces its start)
mkdir -p gpurun_out
for sch in 2,2,2,2,2,2,2,2,2,2 4,4,4,4,4 2,2,4,4,4,4 2,2,2,2,4,4,4 1,1,2,4,4,4,4 3,3,3,3,4,4 5,5,5,5 2,3,5,5,5; do
  python bench.py --no-train --no-cpu-baseline --steps 20 --warmup 5 --e2e-schedule $sch 2>gpurun_out/e.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$sch', 'dev', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), 'bf16-host', round(d['e2e_bf16_host']['ms_per_step'], 3), 'slots', d['e2e']['launches_in_flight'])"
done
SC_BENCH_VERBOSE=1 python bench.py --config train --no-cpu-baseline --steps 20 --warmup 5 2>&1 >/dev/null | grep "train gemm" | sort -k9 -n -r | head -50
