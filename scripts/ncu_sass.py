"""Top stall SASS instructions of an ncu report: python scripts/ncu_sass.py rep.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; items = []
for i, r in enumerate(rows):
    if "Address" in r and "Source" in r:
        hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    try: v = float(d["Warp Stall Sampling (All Samples)"])
    except Exception: continue
    reasons = {k: float(d[k]) for k in hdr if k.startswith("stall_") and "(Not" not in k and d[k] not in ("", "0")}
    items.append((v, len(items), d["Source"].strip(), reasons, d["Instructions Executed"]))
tot = sum(x[0] for x in items) or 1
print("total samples", tot, "instructions", len(items))
for v, idx, src, reasons, ex in sorted(items, key=lambda x: -x[0])[:top]:
    rs = ",".join(f"{k[6:]}={int(x)}" for k, x in sorted(reasons.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*v/tot:5.1f}%  #{idx:5d} ex={ex:>7} {src[:70]:70s} {rs}")
