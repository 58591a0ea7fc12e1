"""Hot CUDA source lines of one kernel in an ncu report: python scripts/ncu_src.py rep.ncu-rep <launch index> [top N]
SASS-level stall samples aggregated per CUDA source line (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys, io, collections
rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fn, cur_file, hdr, col = "?", "?", None, None
cur_line, cur_src = "", ""
seen = set()
agg = collections.defaultdict(lambda: [0, collections.Counter(), "", 0])
tot = 0
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        fn = r[1]; continue
    if r[0] == "Line No":
        hdr = r
        col = {}
        for i, n in enumerate(hdr):
            col.setdefault(n, i)
        src_cols = [i for i, n in enumerate(hdr) if n == "Source"]
        continue
    if hdr is None or len(r) < len(hdr) or "# Samples" not in col:
        continue
    if r[col["Line No"]]:
        cur_line, cur_src = r[col["Line No"]], r[src_cols[0]].strip()   # a CUDA line; its SASS rows follow
        continue
    addr = r[col["Address"]]
    if not addr or (addr, cur_file, cur_line) in seen:
        continue
    seen.add((addr, cur_file, cur_line))
    try:
        n = int(r[col["# Samples"]])
    except ValueError:
        continue
    tot += n
    key = (cur_file, cur_line)
    a = agg[key]
    a[0] += n
    a[3] += int(r[col["Instructions Executed"]] or 0)
    a[2] = cur_src
    for s in hdr:
        if s.startswith("stall_") and "Not Issued" not in s:
            v = int(r[col[s]] or 0)
            if v:
                a[1][s[6:]] += v
print("function:", fn[:200])
print("total samples", tot)
for (f, ln), (n, st, src, ninst) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n:7d} {100*n/max(tot,1):5.1f}% inst={ninst:9d} {f[:20]:20s}:{ln:>5s} {src[:95]:95s} {' '.join(f'{k}={v}' for k, v in st.most_common(3))}")
