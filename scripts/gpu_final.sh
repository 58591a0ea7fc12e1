#!/bin/bash
# what the driver runs at round end: gpu tests, smoke, reference arm, default bench at N=1 and N=2
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu -x > gpurun_out/tests.log 2>&1; echo "tests exit=$? $(tail -n 1 gpurun_out/tests.log)"
grep -E "^E  |Error|FAILED" gpurun_out/tests.log | head -20
timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -1
( time timeout -s KILL 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real; echo "ref exit=$?"; head -c 600 gpurun_out/bench_ref.json; echo
( time timeout -s KILL 900 python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err ) 2>&1 | grep real; echo "bench1 exit=$?"; tail -n 3 gpurun_out/bench1.err
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
( time timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > gpurun_out/bench2.json 2> gpurun_out/bench2.err ) 2>&1 | grep real; echo "bench2 exit=$?"; tail -n 3 gpurun_out/bench2.err
fi
python - <<'PY'
import json
for f in ("gpurun_out/bench1.json","gpurun_out/bench2.json"):
    try:
        txt=open(f).read().strip().splitlines()
        print(f, "stdout lines:", len(txt))
        d=json.loads(txt[-1]); t=d["train"]
        print("  infer", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "clocks", d["clocks"])
        r=d["roofline"]; print("  roofline achieved", round(r["achieved"],1), "frac", round(r["frac"],3), "traffic", r["traffic"], "share", round(r["share_of_step"],3), "dom", r["dominant_shape"])
        print("  cpu", d["cpu_baseline"])
        print("  train img/s", round(t["value"]), "ms", round(t["ms_per_step"],3), "launches", t["gpu_launches_per_step"], "roof", t["roofline"])
    except Exception as e: print(f, "ERR", e)
PY
