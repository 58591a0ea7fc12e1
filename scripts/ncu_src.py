"""Summarise an ncu report's CUDA-source hot lines: python scripts/ncu_src.py report.ncu-rep [kernel-regex] [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else "."; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.defaultdict(float)
hdr = None
for r in rows:
    if "Source" in r and "# Samples" in r:
        hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    try: v = float(d["Warp Stall Sampling (All Samples)"])
    except Exception: continue
    agg[(d.get("#", d.get("Address", "")), d["Source"].strip()[:120])] += v
tot = sum(agg.values()) or 1
for (ln, src), v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{100*v/tot:5.1f}%  L{ln:>5}  {src}")
