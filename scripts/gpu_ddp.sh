#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/ddp_check.py 2>&1 | grep -E "rank|Error|error" | head -20
for ex in allreduce sharded; do
  timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --no-cpu-baseline --exchange $ex > gpurun_out/b2_$ex.json 2> gpurun_out/b2_$ex.err
  python -c "import json;d=json.loads(open('gpurun_out/b2_$ex.json').read().strip().splitlines()[-1]);t=d['train'];print('$ex', 'train img/s', round(t['value']), 'ms', round(t['ms_per_step'],3), 'infer', round(d['value']))" || tail -5 gpurun_out/b2_$ex.err
done
timeout -s KILL 600 python bench.py --no-cpu-baseline > gpurun_out/b1.json 2> gpurun_out/b1.err
python -c "import json;d=json.loads(open('gpurun_out/b1.json').read().strip().splitlines()[-1]);t=d['train'];print('1gpu', 'train img/s', round(t['value']), 'ms', round(t['ms_per_step'],3), 'infer', round(d['value']))"
