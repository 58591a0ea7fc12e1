#!/bin/bash
# same-box A/B of two builds of the library (SC_LIB_PATH = csrc/libsc_b200_base.so vs the in-tree build): kernel timings + bench
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -q -x 2>&1 | tail -3
for v in base new base new; do
  if [ $v = base ]; then export SC_LIB_PATH=$PWD/sparse-image-captioning_b200/csrc/libsc_b200_base.so; else unset SC_LIB_PATH; fi
  echo "== $v"
  timeout -s KILL 300 python scripts/dec_kernels.py --hints 20003256 2>&1 | grep -E "gemm|topk|cross|self_attn_step t=8" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(f\"   {d['kernel']:45s} {d['us']:8.2f} us\")"
  python bench.py --no-cpu-baseline --no-train --steps 20 --warmup 5 2> gpurun_out/ab_$v.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(f\"   bench dev {d['ms_per_step']:.3f} ms  e2e {d['e2e']['ms_per_step']:.3f} ms\")"
done
