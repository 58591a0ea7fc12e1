"""Time the tcgen05 GEMM on the SMP-training shapes with their real epilogues (fp32 residual stream, dropout, bf16 / fp32
stores) under each (tile_n, stages) config, inside a CUDA graph.  Diagnostic for the tile-selection heuristic."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_caption_b200.kernels as K
dev = "cuda"
K.set_pdl(int(os.environ.get("SC_PDL_MASK", "0")))
# (name, M, N, K, out dtype, residual, relu, dropout)
cases = []
for M in (4250, 1800):
    cases += [("o/ff-res", M, 512, 512, torch.float32, True, False, 0.1), ("qkv", M, 1536, 512, torch.bfloat16, False, False, 0.0),
              ("ff1", M, 2048, 512, torch.bfloat16, False, True, 0.1), ("ff2", M, 512, 2048, torch.float32, True, False, 0.1),
              ("dx512", M, 512, 512, torch.float32, False, False, 0.0), ("dx1536", M, 512, 1536, torch.float32, False, False, 0.0),
              ("dx-ff2", M, 2048, 512, torch.float32, False, False, 0.0), ("dx-ff1", M, 512, 2048, torch.float32, False, False, 0.0)]
cases += [("cq", 4250, 512, 512, torch.bfloat16, False, False, 0.0), ("ckv", 1800, 1024, 512, torch.bfloat16, False, False, 0.0),
          ("gen", 4250, 10000, 512, torch.float32, False, False, 0.0), ("dx-gen", 4250, 512, 10000, torch.float32, False, False, 0.0),
          ("att_embed", 1800, 512, 2048, torch.float32, False, True, 0.5)]
cfgs = [(0, 0), (64, 3), (64, 4), (64, 6), (128, 3), (128, 5), (256, 3)]
for name, M, N, Kd, odt, has_res, relu, p in cases:
    x = torch.randn(M, Kd, device=dev).bfloat16(); w = torch.randn(N, Kd, device=dev).bfloat16()
    b = torch.randn(N, device=dev) if not name.startswith("dx") else None
    res = torch.randn(M, N, device=dev) if has_res else None
    outs = [torch.empty(M, N, device=dev, dtype=odt) for _ in range(4)]
    line = []
    for bn, st in cfgs:
        tile = st * 1000 + bn
        try:
            def run(i):
                if p > 0:
                    K.linear_dropout(x, w, b, residual=res, relu=relu, out=outs[i % 4], p=p, drop_seed=5, drop_stream=7, tile_n=tile)
                else:
                    K.linear(x, w, b, residual=res, relu=relu, out=outs[i % 4], tile_n=tile)
            run(0); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(40): run(i)
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 80
            line.append(f"{bn}x{st}:{us:6.1f}")
        except Exception as ex:
            line.append(f"{bn}x{st}:  ERR ")
    print(f"{name:10s} M={M:5d} N={N:5d} K={Kd:5d}  " + "  ".join(line), flush=True)
