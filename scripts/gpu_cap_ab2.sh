#!/bin/bash
# persistent-GEMM grid cap x batching (SC_GEMM_MAX_CTAS): GEMMs of several launches share the SMs, attention kernels take the rest
run() {
  cap=$1; shift
  SC_GEMM_MAX_CTAS=$cap python bench.py --no-cpu-baseline --no-train --steps 20 --warmup 5 "$@" 2> gpurun_out/e.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(f\"cap=$cap $*  dev {d['ms_per_step']:.3f} ms  e2e {d['e2e']['ms_per_step']:.3f} ms\")"
}
run 40 --coalesce 5 --slots 4
run 32 --coalesce 5 --slots 4
run 24 --coalesce 5 --slots 4
run 48 --coalesce 4 --slots 5
run 32 --coalesce 4 --slots 5
run 48 --coalesce 2 --slots 10
run 32 --coalesce 2 --slots 10
run 24 --coalesce 2 --slots 10
run 0 --coalesce 5 --slots 4
