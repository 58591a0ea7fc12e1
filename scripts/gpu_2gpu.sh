#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_trainer_gpu.py tests/test_dropin_gpu.py -q -m gpu > gpurun_out/t_train.log 2>&1; echo "tests exit=$? $(tail -n 1 gpurun_out/t_train.log)"
grep -E "^E  |Error|FAILED" gpurun_out/t_train.log | head -20
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench2 exit=$?"
tail -c 3000 gpurun_out/bench2.json; tail -n 8 gpurun_out/bench2.err
timeout -s KILL 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench1 exit=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench1.json","gpurun_out/bench2.json"):
    try:
        d=json.load(open(f)); t=d["train"]
        print(f, "infer", round(d["value"]), "e2e", round(d["e2e"]["value"]), "train img/s", round(t["value"]), "ms", round(t["ms_per_step"],3), "launches", t["gpu_launches_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
