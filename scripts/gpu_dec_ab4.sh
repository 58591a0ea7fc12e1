#!/bin/bash
# Decode-step A/B at the driver's own K / W: dynamic batching x slots x tile hints.
mkdir -p gpurun_out
run() {
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py --no-train --no-cpu-baseline --steps 20 --warmup 5 "$@" > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/ab_{n}.json"))
    r=d["roofline"]
    print(f"{n:28s} dev {d['ms_per_step']:.3f} ms  e2e {d['e2e']['ms_per_step']:.3f} ms  gemm alone {r['achieved']:.0f} TF/s in-flight {r['achieved_in_flight']:.0f} TF/s  slots {d['config']['batches_in_flight']}")
except Exception as ex:
    print(n, "FAILED", ex, open(f"gpurun_out/ab_{n}.err").read()[-400:])
PY
}
T=qkv=3256,o=3256,cq=3256,co=3256,ff1=3256,ff2=3256
run k20_base A=1 --
run k20_g5_s2 A=1 -- --coalesce 5 --slots 2
run k20_g5_s2_t SC_DEC_TILES=$T -- --coalesce 5 --slots 2
run k20_g5_s2_tpc SC_GEMM_TPC=2 SC_DEC_TILES=$T -- --coalesce 5 --slots 2
run k20_g5_s4 A=1 -- --coalesce 5 --slots 4
run k20_g4_s5 A=1 -- --coalesce 4 --slots 5
run k20_g2_s5_tpc SC_GEMM_TPC=2 SC_DEC_TILES=$T -- --coalesce 2 --slots 5
run k20_g10_s2 A=1 -- --coalesce 10 --slots 2
