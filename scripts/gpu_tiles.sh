#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "topk or partials" 2>&1 | tail -1
run() { SC_DEC_TILES="$1" timeout -s KILL 300 python bench.py --steps 16 --warmup 8 --slots $2 --no-train --no-cpu-baseline > gpurun_out/bm.json 2> gpurun_out/bm.err; python -c "import json;d=json.load(open('gpurun_out/bm.json'));print('tiles=[$1] slots=$2', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))" || tail -3 gpurun_out/bm.err; }
run "" 4
run "o=5128,co=5128,ff2=5128" 4
run "o=5128,co=5128,ff2=5128" 6
run "" 8
run "o=3256,co=3256,cq=3128,ff2=3256" 8
run "o=5128,co=5128,cq=3128,ff2=5128" 8
