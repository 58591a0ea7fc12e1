#!/bin/bash
# 8-GPU run of the default bench (inference + SMP training arm) with NCCL's own description of what it set up
# (gpurun --gpus 8 -- 'bash scripts/gpu_n8.sh'); NCCL lines -> gpurun_out/nccl_n8.txt, bench line -> gpurun_out/bench_n8.json
mkdir -p gpurun_out
N=${N:-8}
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,ENV,TUNING timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "exit=$?"
grep -E "NCCL INFO" gpurun_out/bench_n$N.err | grep -E "NVLS|Ring|Tree|hannels|Connected all|comm .* rank 0 |Algo|algo|P2P|nvls" | grep -E "\[0\]|rank 0" | sed 's/^.*NCCL INFO //' | sort | uniq -c | sort -rn | head -40 > gpurun_out/nccl_n$N.txt
cat gpurun_out/nccl_n$N.txt | head -30
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1]); t = d["train"]
print("infer", round(d["value"]), "captions/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["ms_per_step"], 3), "bf16-host", round(d["e2e_bf16_host"]["ms_per_step"], 3))
print("train", round(t["value"]), "images/s", round(t["ms_per_step"], 3), "ms", t.get("collective", "")[:80])
PY
