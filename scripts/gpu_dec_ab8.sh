#!/bin/bash
# do the decode GEMMs leave room on an SM for the HBM-bound kernels of the other launches?  small-footprint tile configurations
# (128-wide, 3 stages, 4 epilogue warps: 113 KB smem, 256 threads x 128 registers -> two CTAs / other kernels share the SM)
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
S128=qkv=3128,o=3128,cq=3128,co=3128,ff1=3128,ff2=3128
S64=qkv=3064,o=3064,cq=3064,co=3064,ff1=3064,ff2=3064
MIX=qkv=20003256,o=3128,cq=3128,co=3128,ff1=20003256,ff2=3128
run base_g5s4 A=1 -- --coalesce 5 --slots 4
run s128_g5s4 SC_DEC_TILES=$S128 -- --coalesce 5 --slots 4
run s128_g5s8 SC_DEC_TILES=$S128 -- --coalesce 5 --slots 8
run s128_g2s10 SC_DEC_TILES=$S128 -- --coalesce 2 --slots 10
run s128x2_g5s4 SC_DEC_TILES=$S128 SC_GEMM_PER_SM=2 -- --coalesce 5 --slots 4
run mix_g5s4 SC_DEC_TILES=$MIX -- --coalesce 5 --slots 4
run s64_g5s4 SC_DEC_TILES=$S64 -- --coalesce 5 --slots 4
