#!/bin/bash
# dynamic batching x pipeline slots after the session-c kernel changes (driver K / W)
mkdir -p gpurun_out
run() {
  name=$1; shift
  python bench.py --no-train --no-cpu-baseline --steps 20 --warmup 5 "$@" > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/ab_{n}.json"))
    r=d["roofline"]
    print(f"{n:16s} dev {d['ms_per_step']:.3f} ms  e2e {d['e2e']['ms_per_step']:.3f} ms  bf16-host {d['e2e_bf16_host']['ms_per_step']:.3f}  gemm alone {r['achieved']:.0f} TF/s in-flight {r['achieved_in_flight']:.0f} TF/s")
except Exception as ex:
    print(n, "FAILED", ex, open(f"gpurun_out/ab_{n}.err").read()[-300:])
PY
}
run g5_s4 --coalesce 5 --slots 4
run g5_s2 --coalesce 5 --slots 2
run g5_s3 --coalesce 5 --slots 3
run g5_s6 --coalesce 5 --slots 6
run g4_s5 --coalesce 4 --slots 5
run g10_s2 --coalesce 10 --slots 2
run g10_s3 --coalesce 10 --slots 3
run g2_s10 --coalesce 2 --slots 10
run g5_s4_e4 --coalesce 5 --slots 4 --e2e-coalesce 4 --e2e-slots 5
run g5_s4_e1 --coalesce 5 --slots 4 --e2e-coalesce 1 --e2e-slots 10
