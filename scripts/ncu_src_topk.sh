#!/bin/bash
mkdir -p gpurun_out
SC_NCU_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:sc_gemm_bf16_kernel<.int.256, .bool.0, .int.3, .int.5" -s 3 -c 1 -f -o gpurun_out/topk_src python scripts/profile_step.py 512 dense > gpurun_out/ncu_c.log 2>&1
ncu -i gpurun_out/topk_src.ncu-rep --page source --csv --print-source cuda > gpurun_out/topk_src.csv 2> gpurun_out/topk_src.err
head -c 1500 gpurun_out/topk_src.csv; echo; wc -l gpurun_out/topk_src.csv; tail -3 gpurun_out/topk_src.err
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/topk_src.csv")))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r:
        hdr = r; start = i + 1; break
print("header:", hdr)
if hdr:
    col = [c for c in hdr if "Sampl" in c]
    print("sampling columns:", col)
    key = col[0] if col else None
    agg = collections.defaultdict(float)
    for r in rows[start:]:
        if len(r) < len(hdr): continue
        d = dict(zip(hdr, r))
        try: v = float(d[key])
        except Exception: continue
        agg[(d.get("#", ""), d["Source"].strip()[:110])] += v
    tot = sum(agg.values()) or 1
    for (ln, src), v in sorted(agg.items(), key=lambda kv: -kv[1])[:35]:
        print(f"{100*v/tot:5.1f}%  L{ln:>5}  {src}")
PY
rm -f gpurun_out/*.ncu-rep gpurun_out/topk_src.csv
