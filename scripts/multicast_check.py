"""GEMM variants (SC_GEMM_MULTICAST = 0 no clusters, 1 multicast B, 2 default: CTA pairs for >= 64 M blocks, 3 pairs everywhere): correctness against torch, in-graph timing, cuBLAS beside it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_caption_b200.kernels as K
dev = "cuda"
TILE = int(os.environ.get("TILE", "3256"))
shapes = [(18432, 1536, 512), (18432, 2048, 512), (18432, 512, 2048), (18432, 1024, 512), (1536, 10000, 512), (4250, 10000, 512), (2000, 520, 136)]
for (M, N, Kd) in shapes:
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, Kd, generator=g).bfloat16().to(dev); w = (torch.randn(N, Kd, generator=g) * 0.1).bfloat16().to(dev); b = torch.randn(N, generator=g).to(dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    K.linear(x, w, b, out=out, tile_n=TILE)
    torch.cuda.synchronize()
    ref = (x.float() @ w.float().t() + b)
    err = float((out.float() - ref).abs().max() / ref.abs().max())
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(20): K.linear(x, w, b, out=out, tile_n=TILE)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    gr2 = torch.cuda.CUDAGraph()
    o2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    torch.matmul(x, w.t(), out=o2); torch.cuda.synchronize()
    with torch.cuda.graph(gr2):
        for i in range(20): torch.matmul(x, w.t(), out=o2)
    gr2.replay(); torch.cuda.synchronize()
    e0.record(); gr2.replay(); e1.record(); torch.cuda.synchronize()
    us2 = e0.elapsed_time(e1) * 1e3 / 20
    print(f"cuBLAS (no bias): {us2:7.1f} us  {2*M*N*Kd/us2/1e6:6.0f} TF/s")
    print(f"TILE={TILE} MULTICAST={os.environ.get('SC_GEMM_MULTICAST','default')} M={M:5d} N={N:5d} K={Kd:4d}: {us:7.1f} us  {2*M*N*Kd/us/1e6:6.0f} TF/s  rel err {err:.2e}", flush=True)
