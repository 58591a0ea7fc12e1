"""One launch of every GEMM shape of configs[2] (after a warm-up launch and an L2 flush) - the target of scripts/ncu_gemm_traffic.sh.
Shapes: the decode GEMMs at 1536 rows (one batch) and 7680 rows (5 coalesced batches, the bench default) with the engine's tile
hints, and the encoder GEMMs of the coalesced batch (92160 rows)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_caption_b200.kernels as K
dev = "cuda"
DEC = [  # (N, K, out bytes, residual, relu)
    (1536, 512, 2, 0, 0), (512, 512, 4, 1, 0), (512, 512, 2, 0, 0), (2048, 512, 2, 0, 1), (512, 2048, 4, 1, 0), (10000, 512, 0, 0, 0)]
ENC = [(512, 2048, 4, 0, 1), (1536, 512, 2, 0, 0), (512, 512, 4, 1, 0), (2048, 512, 2, 0, 1), (512, 2048, 4, 1, 0), (1024, 512, 2, 0, 0)]
SHAPES = [(1536,) + s + (0,) for s in DEC] + [(7680,) + s + (20003256,) for s in DEC] + [(92160,) + s + (0,) for s in ENC]
if __name__ == "__main__":
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for M, N, Kd, ys, res, relu, tile in SHAPES:
        x = torch.randn(M, Kd, device=dev).bfloat16(); w = torch.randn(N, Kd, device=dev).bfloat16(); b = torch.randn(N, device=dev)
        if ys == 0:
            part = torch.empty(M, K.linear_topk_parts(N), 12, device=dev)
            run = lambda: K.linear_topk(x, w, b, part, candidates=3)
        else:
            y = torch.randn(M, N, device=dev).to(torch.bfloat16 if ys == 2 else torch.float32)
            run = lambda: K.linear(x, w, b, residual=y if res else None, relu=bool(relu), out=y, tile_n=tile)   # residual GEMMs run in place
        run(); torch.cuda.synchronize()
        flush.zero_(); torch.cuda.synchronize()          # evict the operands: the captured launch reads them from HBM
        run(); torch.cuda.synchronize()
