"""SC_GEMM_MULTICAST=1: 2-CTA clusters with a multicast B tile - correctness against torch and in-graph timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_caption_b200.kernels as K
dev = "cuda"
shapes = [(18432, 1536, 512), (18432, 2048, 512), (18432, 512, 2048), (18432, 1024, 512), (1536, 10000, 512), (4250, 10000, 512), (2000, 520, 136)]
for (M, N, Kd) in shapes:
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, Kd, generator=g).bfloat16().to(dev); w = (torch.randn(N, Kd, generator=g) * 0.1).bfloat16().to(dev); b = torch.randn(N, generator=g).to(dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    K.linear(x, w, b, out=out, tile_n=3256)
    torch.cuda.synchronize()
    ref = (x.float() @ w.float().t() + b)
    err = float((out.float() - ref).abs().max() / ref.abs().max())
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(20): K.linear(x, w, b, out=out, tile_n=3256)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    print(f"MULTICAST={os.environ.get('SC_GEMM_MULTICAST','0')} M={M:5d} N={N:5d} K={Kd:4d}: {us:7.1f} us  {2*M*N*Kd/us/1e6:6.0f} TF/s  rel err {err:.2e}", flush=True)
