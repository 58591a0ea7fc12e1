#!/bin/bash
# larger coalesced device batches x tiles per persistent GEMM CTA (driver K / W)
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
    print(f"{n:16s} dev {d['ms_per_step']:.3f} ms  e2e {d['e2e']['ms_per_step']:.3f} ms  gemm alone {r['achieved']:.0f} TF/s in-flight {r['achieved_in_flight']:.0f} TF/s")
except Exception as ex:
    print(n, "FAILED", ex, open(f"gpurun_out/ab_{n}.err").read()[-300:])
PY
}
T3=qkv=30003256,o=30003256,cq=30003256,co=30003256,ff1=30003256,ff2=30003256
T4=qkv=40003256,o=40003256,cq=40003256,co=40003256,ff1=40003256,ff2=40003256
run g5_s4 A=1 -- --coalesce 5 --slots 4
run g10_s2 A=1 -- --coalesce 10 --slots 2
run g10_s2_t3 SC_DEC_TILES=$T3 -- --coalesce 10 --slots 2
run g10_s2_t4 SC_DEC_TILES=$T4 -- --coalesce 10 --slots 2
run g10_s3_t3 SC_DEC_TILES=$T3 -- --coalesce 10 --slots 3
run g20_s1_t4 SC_DEC_TILES=$T4 -- --coalesce 20 --slots 1
run g20_s2_t4 SC_DEC_TILES=$T4 -- --coalesce 20 --slots 2
run g5_s4_t3 SC_DEC_TILES=$T3 -- --coalesce 5 --slots 4
