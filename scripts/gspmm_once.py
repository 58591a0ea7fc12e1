"""A few launches of sc_gspmm (and the dense GEMM) at one decode shape - ncu target (scripts/ncu_gspmm.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_caption_b200.kernels as K
dev = "cuda"
M, N, Kd = 1536, int(os.environ.get("SC_N", "512")), int(os.environ.get("SC_K", "512"))
x = torch.randn(M, Kd, device=dev).bfloat16(); b = torch.randn(N, device=dev); y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
for sp in (0.991, 0.95):
    w = torch.randn(N, Kd, device=dev); w[torch.rand(N, Kd, device=dev) < sp] = 0
    gw = K.GsWeight(w.bfloat16().float())
    for _ in range(3):
        K.gspmm(x, gw, b, out=y)
    torch.cuda.synchronize()
wd = torch.randn(N, Kd, device=dev).bfloat16()
for _ in range(3):
    K.linear(x, wd, b, out=y)
torch.cuda.synchronize()
