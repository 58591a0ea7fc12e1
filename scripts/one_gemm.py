"""Runs one GEMM variant a few times (for ncu --set full captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_caption_b200.kernels as K
M, N, Kd, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
dev = "cuda"
x = torch.randn(M, Kd, device=dev).bfloat16(); w = torch.randn(N, Kd, device=dev).bfloat16(); b = torch.randn(N, device=dev)
y32 = torch.randn(M, N, device=dev); yb = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
st_in = torch.rand(M, Kd // 32, 2, device=dev) + 1; lnc = torch.randn(N, device=dev); st_out = torch.empty(M, max(N // 32, 1), 2, device=dev)
for _ in range(5):
    if mode == "plain": K.linear(x, w, b, out=yb)
    elif mode == "res": K.linear(x, w, b, residual=y32, out=y32)
    elif mode == "ln": K.linear_ln(x, w, b, out=yb, ln_stats=st_in, ln_c=lnc)
    elif mode == "produce": K.linear_ln(x, w, b, residual=y32, out=y32, out_bf16=yb, stats_out=st_out)
torch.cuda.synchronize()
