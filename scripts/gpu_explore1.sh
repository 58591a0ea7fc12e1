#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python scripts/gemm_sweep.py > gpurun_out/gemm_sweep.txt 2>&1; echo "sweep exit=$?"
for s in 1 2 8 12; do
  timeout -s KILL 300 python bench.py --steps 12 --warmup 12 --slots $s --no-train --no-cpu-baseline > gpurun_out/bench_s$s.json 2> gpurun_out/bench_s$s.err; echo "slots $s exit=$?"
  python -c "import json;d=json.load(open('gpurun_out/bench_s$s.json'));print($s, d['ms_per_step'], d['e2e']['ms_per_step'])"
done
timeout -s KILL 300 python scripts/profile_train.py > gpurun_out/profile_train.txt 2>&1; echo "ptrain exit=$?"
cat gpurun_out/gemm_sweep.txt; cat gpurun_out/profile_train.txt
