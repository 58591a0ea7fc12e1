#!/bin/bash
# Decode-step A/B on one box: pipeline slots, CTA-pair GEMMs (cta_group::2, B tile split across the pair), tile hints, LN fold.
mkdir -p gpurun_out
run() { # name, env..., -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py --no-train --no-cpu-baseline --steps 40 --warmup 8 "$@" > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
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
run base8 A=1 -- --slots 8
run base5 A=1 -- --slots 5
run base10 A=1 -- --slots 10
run pairs8 SC_GEMM_MULTICAST=3 SC_DEC_TILES=qkv=3256,o=3256,cq=3256,co=3256,ff1=3256,ff2=3256 -- --slots 8
run all256_8 SC_DEC_TILES=qkv=3256,o=3256,cq=3256,co=3256,ff1=3256,ff2=3256 -- --slots 8
run lnfold8 A=1 -- --slots 8 --ln-fold
