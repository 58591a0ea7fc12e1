"""K3a vs K3b vs K3b' at decode sizes: dense tcgen05 GEMM (pre-masked bf16 weights), CSR SpMM and sliced-ELL SpMM for the
decoder's linear shapes at 80-99.1 % unstructured sparsity, timed inside a CUDA graph (40 back-to-back launches over
rotating outputs).  Output: one table row per (shape, sparsity); DESIGN.md quotes it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_caption_b200.kernels as K
dev = "cuda"
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1536
def timeit(fn, reps=40):
    fn(0); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps): fn(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (2 * reps)
print(f"rows M={R}; microseconds per launch (in-graph)")
print(f"{'shape NxK':>12s} {'sparsity':>8s} {'dense':>8s} {'csr':>8s} {'sell':>8s}  {'nnz':>8s} {'sell pad':>8s}")
torch.manual_seed(0)
for (Nn, Kd) in [(512, 512), (1536, 512), (2048, 512), (512, 2048), (10000, 512)]:
    x = torch.randn(R, Kd, device=dev).bfloat16(); b = torch.randn(Nn, device=dev)
    ys = [torch.empty(R, Nn, device=dev, dtype=torch.bfloat16) for _ in range(4)]
    for sp in (0.8, 0.9, 0.95, 0.99, 0.991):
        w = torch.randn(Nn, Kd, device=dev)
        keep = torch.rand(Nn, Kd, device=dev) >= sp
        w = (w * keep).bfloat16()
        t_dense = timeit(lambda i: K.linear(x, w, b, out=ys[i % 4]))
        csr = K.CsrWeight(w.float(), torch.bfloat16)
        t_csr = timeit(lambda i: K.csr_spmm(x, csr, b, out=ys[i % 4]), reps=10) if sp >= 0.9 else float("nan")
        sw = K.SellWeight(w.float(), torch.bfloat16)
        t_sell = timeit(lambda i: K.sell_spmm(x, sw, b, out=ys[i % 4]))
        ref = ys[0].float().clone(); K.linear(x, w, b, out=ys[1]); err = float((ys[1].float() - ref).abs().max() / ref.abs().max())
        print(f"{Nn:>6d}x{Kd:<5d} {sp:8.3f} {t_dense:8.1f} {t_csr:8.1f} {t_sell:8.1f}  {sw.nnz:8d} {sw.padded:8d}  relerr(sell vs dense) {err:.1e}", flush=True)
