#!/bin/bash
# SMP training arm at N GPUs under different NCCL CTA budgets (the NVLS all-reduce runs 24 channels = 24 CTAs by default)
mkdir -p gpurun_out
N=${N:-8}
for cfg in "default" "NCCL_MAX_CTAS=16" "NCCL_MAX_CTAS=8" "NCCL_NVLS_NCHANNELS=8"; do
  envs=(); [ "$cfg" != "default" ] && envs=("$cfg")
  env "${envs[@]}" timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --config train --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/train_n${N}_${cfg//=/_}.json 2> gpurun_out/train_n${N}.err
  python - "$cfg" "gpurun_out/train_n${N}_${cfg//=/_}.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:24s} train {d['ms_per_step']:.3f} ms/step  {d['value']:.0f} images/s")
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
