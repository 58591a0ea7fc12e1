"""Dense tcgen05 GEMM vs sliced-ELL vs gather SpMM (sc_gspmm) at decode sizes, per sparsity level: us per launch inside a CUDA
graph, alone and with 8 streams running the same kernel concurrently (the timed region's regime); effective TF/s = 2 M nnz / t."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_caption_b200.kernels as K
dev = "cuda"
M = int(os.environ.get("SC_M", "1536"))
SHAPES = [(512, 512), (1536, 512), (2048, 512), (512, 2048), (10000, 512)]
LEVELS = [0.80, 0.90, 0.95, 0.975, 0.9875, 0.991]
def time_graph(run, streams=1):
    sts = [torch.cuda.Stream() for _ in range(streams)]
    graphs = []
    for st in sts:
        run(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20): run()
        graphs.append(g)
    cur = torch.cuda.current_stream()
    def go():
        for st, g in zip(sts, graphs):
            st.wait_stream(cur)
            with torch.cuda.stream(st): g.replay(); g.replay()
        for st in sts: cur.wait_stream(st)
    go(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); go(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (streams * 40)
print(f"M = {M} rows; us per launch alone / 8 concurrent streams (wall per launch); eff = 2*M*nnz/t of the gather SpMM alone")
for N, Kd in SHAPES:
    x = torch.randn(M, Kd, device=dev).bfloat16(); b = torch.randn(N, device=dev)
    y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    wd = torch.randn(N, Kd, device=dev).bfloat16()
    d1, d8 = time_graph(lambda: K.linear(x, wd, b, out=y)), time_graph(lambda: K.linear(x, wd, b, out=y), 8)
    print(f"--- N={N} K={Kd}: dense tcgen05 {d1:6.2f} / {d8:6.2f}")
    for sp in LEVELS:
        w = torch.randn(N, Kd, device=dev)
        w[torch.rand(N, Kd, device=dev) < sp] = 0
        gw = K.GsWeight(w.bfloat16().float())
        g1, g8 = time_graph(lambda: K.gspmm(x, gw, b, out=y)), time_graph(lambda: K.gspmm(x, gw, b, out=y), 8)
        line = f"    sparsity {sp:6.4f} nnz {gw.nnz:8d}: gather {g1:6.2f} / {g8:6.2f}  ({2.0 * M * gw.nnz / g1 / 1e6:5.1f} TF/s eff)"
        if Kd * 32 <= 200 * 1024 and N <= 2048:
            sw = K.SellWeight(w.bfloat16().float(), torch.bfloat16)
            s1 = time_graph(lambda: K.sell_spmm(x, sw, b, out=y))
            line += f"   sliced-ELL {s1:6.2f}"
        win = "gather" if g8 < d8 else "dense"
        print(line + f"   -> winner in flight: {win} ({min(g8, d8) / max(g8, d8):.2f})", flush=True)
