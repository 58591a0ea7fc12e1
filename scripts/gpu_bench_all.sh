#!/bin/bash
# Every bench configuration at the driver's K / W on one box (+ the reference arms) -> gpurun_out/bench_<config>.json
mkdir -p gpurun_out
for c in infer train acort scst; do
  python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err || tail -5 gpurun_out/bench_$c.err
  python bench.py --config $c --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$c.json 2> gpurun_out/bench_ref_$c.err || tail -5 gpurun_out/bench_ref_$c.err
done
python - <<'PY'
import json
for c in ("infer", "train", "acort", "scst"):
    try:
        d = json.load(open(f"gpurun_out/bench_{c}.json")); r = json.load(open(f"gpurun_out/bench_ref_{c}.json"))
        rf = d["roofline"]
        print(f"{c:6s} {d['metric']:34s} value {d['value']:10.1f} {d['unit']:11s} {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value']:10.1f}  "
              f"ref {r['value']:.1f}  roofline {rf['achieved']:.0f} TF/s frac {rf['frac']:.3f}" + (f" in-flight {rf.get('frac_in_flight') or 0:.3f}" if 'frac_in_flight' in rf else ""))
        if d.get("train"): print("       train:", round(d["train"]["value"]), "img/s", round(d["train"]["ms_per_step"], 3), "ms")
        if rf.get("hbm_kernels"):
            for k in rf["hbm_kernels"]: print(f"       {k['kernel']:36s} {k['us_per_launch']:.2f} us  {k['achieved']:.0f} GB/s  frac {k['frac']:.2f}")
    except Exception as ex:
        print(c, "FAILED", ex)
PY
