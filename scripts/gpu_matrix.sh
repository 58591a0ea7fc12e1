#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py tests/test_train_kernels_gpu.py tests/test_engine_gpu.py -q -m gpu -x -k "linear or engine or fold" > gpurun_out/t_linear.log 2>&1; echo "tests exit=$? $(tail -n 1 gpurun_out/t_linear.log)"
grep -E "^E  |Error|FAILED" gpurun_out/t_linear.log | head
timeout -s KILL 300 python scripts/ln_sweep.py 2>&1 | tee gpurun_out/ln_sweep.txt
for flags in "" "--ln-fold"; do
for s in 1 4; do
  timeout -s KILL 300 python bench.py --steps 12 --warmup 12 --slots $s --no-train --no-cpu-baseline $flags > gpurun_out/bm.json 2> gpurun_out/bm.err
  python -c "import json;d=json.load(open('gpurun_out/bm.json'));print('$flags', $s, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))"
done; done
