#!/bin/bash
# Decode-step A/B (round 2, session c) at the driver's K / W: CTA pairs at M = 7680, tiles per persistent CTA
source scripts/gpu_dec_ab_lib.sh
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
A=qkv=3256,o=3256,cq=3256,co=3256,ff1=3256,ff2=3256
T3=qkv=30003256,o=30003256,cq=30003256,co=30003256,ff1=30003256,ff2=30003256
MIX=qkv=20003256,o=3256,cq=3256,co=3256,ff1=20003256,ff2=3256
run c_base A=1 --
run c_pairs SC_GEMM_MULTICAST=3 --
run c_tpc1 SC_DEC_TILES=$A --
run c_tpc3 SC_DEC_TILES=$T3 --
run c_pairs_tpc1 SC_GEMM_MULTICAST=3 SC_DEC_TILES=$A --
run c_mix SC_DEC_TILES=$MIX --
run c_base2 A=1 --
