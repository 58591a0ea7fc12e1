"""fp32-output GEMM (training-logits shape) in-graph timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_caption_b200.kernels as K
for (M, N, Kd) in ((4250, 10000, 512), (1536, 10000, 512)):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(M, Kd, generator=g).bfloat16().cuda(); w = (torch.randn(N, Kd, generator=g) * 0.1).bfloat16().cuda(); b = torch.randn(N, generator=g).cuda()
    out = torch.empty(M, N, device="cuda")
    K.linear(x, w, b, out=out, tile_n=3256); torch.cuda.synchronize()
    err = float((out - (x.float() @ w.float().t() + b)).abs().max())
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(20): K.linear(x, w, b, out=out, tile_n=3256)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    print(f"fp32 out M={M} N={N} K={Kd}: {us:.1f} us {2*M*N*Kd/us/1e6:.0f} TF/s  max err {err:.2e}")
