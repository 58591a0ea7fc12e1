#!/bin/bash
# tiles per persistent CTA of the decode GEMMs (qkv 360 tiles, o / cq / co / ff2 120, ff1 480 at 7680 rows)
run() {
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py --no-cpu-baseline --no-train --steps 20 --warmup 5 "$@" 2> gpurun_out/e.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(f\"$name  dev {d['ms_per_step']:.3f} ms  e2e {d['e2e']['ms_per_step']:.3f} ms  bf16-host {d['e2e_bf16_host']['ms_per_step']:.3f}\")"
}
h() { echo "qkv=$1""0003256,o=$2""0003256,cq=$2""0003256,co=$2""0003256,ff1=$3""0003256,ff2=$2""0003256"; }
run base A=1 --
run d_8_3_10 SC_DEC_TILES=$(h 8 3 10) --
run d_6_2_8 SC_DEC_TILES=$(h 6 2 8) --
run d_12_4_15 SC_DEC_TILES=$(h 12 4 15) --
run d_8_2_10 SC_DEC_TILES=$(h 8 2 10) --
run d_5_3_6 SC_DEC_TILES=$(h 5 3 6) --
run d_8_3_10_g10s2 SC_DEC_TILES=$(h 8 3 10) -- --coalesce 10 --slots 2
run d_8_3_10_g4s5 SC_DEC_TILES=$(h 8 3 10) -- --coalesce 4 --slots 5
