#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -q -m gpu -k "sell or csr or engine" > gpurun_out/t_sell.log 2>&1; echo "tests exit=$? $(tail -n 1 gpurun_out/t_sell.log)"
grep -E "^E  |Error|FAILED" gpurun_out/t_sell.log | head -20
timeout -s KILL 300 python scripts/sparse_sweep.py 2>&1 | tee gpurun_out/sparse_sweep.txt
for be in dense sell; do
  timeout -s KILL 300 python bench.py --steps 12 --warmup 6 --backend $be --no-train --no-cpu-baseline > gpurun_out/bm_$be.json 2> gpurun_out/bm_$be.err
  python -c "import json;d=json.load(open('gpurun_out/bm_$be.json'));print('$be', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['roofline']['kernel_time_breakdown_ms'])"
done
for m in 0 1 2 3; do for r in 1 4; do echo "train pdl=$m ring=$r: $(SC_WALL_ONLY=1 SC_PDL_MASK=$m SC_WGRAD_RING=$r python scripts/profile_train.py 2>&1 | tail -1)"; done; done
