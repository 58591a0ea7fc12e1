"""One launch of every decode GEMM shape of configs[2] (after a warm-up launch) - the target of scripts/ncu_gemm_traffic.sh."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_caption_b200.kernels as K
dev = "cuda"
M = int(os.environ.get("SC_M", "1536"))
SHAPES = [  # (N, K, out bytes, residual, relu)
    (1536, 512, 2, 0, 0), (512, 512, 4, 1, 0), (512, 512, 2, 0, 0), (2048, 512, 2, 0, 1), (512, 2048, 4, 1, 0), (10000, 512, 0, 0, 0)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for N, Kd, ys, res, relu in SHAPES:
    x = torch.randn(M, Kd, device=dev).bfloat16(); w = torch.randn(N, Kd, device=dev).bfloat16(); b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev) if res else None
    if ys == 0:
        part = torch.empty(M, K.linear_topk_parts(N), 12, device=dev)
        run = lambda: K.linear_topk(x, w, b, part, candidates=3)
    else:
        y = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if ys == 2 else torch.float32)
        run = lambda: K.linear(x, w, b, residual=r, relu=bool(relu), out=y)
    run(); torch.cuda.synchronize()
    flush.zero_(); torch.cuda.synchronize()          # evict the operands: the captured launch reads them from HBM
    torch.cuda.nvtx.range_push(f"shape {M},{N},{Kd},{ys},{res}")
    run(); torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
