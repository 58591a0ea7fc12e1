#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
( time timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err ) 2>&1 | grep real
python -c "
import json
d=json.loads(open('gpurun_out/bench_n.json').read().strip().splitlines()[-1]); t=d['train']
print('N', d['n_gpus'], 'infer', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'train img/s', round(t['value']), 'ms', round(t['ms_per_step'],3))
" || tail -20 gpurun_out/bench_n.err
