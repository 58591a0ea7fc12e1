"""Aggregate throughput of the decode-size GEMMs when S independent streams run them concurrently (the engine's pipeline
slots), per tile config: microseconds of GPU time per GEMM = elapsed / (S * launches).  Diagnostic for tile selection."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_caption_b200.kernels as K
dev = "cuda"
R = 1536
cases = [("o/co (res,f32)", 512, 512, True, torch.float32), ("cq (bf16)", 512, 512, False, torch.bfloat16),
         ("qkv (bf16)", 1536, 512, False, torch.bfloat16), ("ff1 (bf16,relu)", 2048, 512, False, torch.bfloat16),
         ("ff2 (res,f32)", 512, 2048, True, torch.float32)]
cfgs = [(0, 0), (64, 3), (64, 6), (128, 3), (128, 5), (256, 3)]
for S in (1, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(S)]
    print(f"--- {S} concurrent streams, M={R}: us of wall time per GEMM (elapsed / (S*40*2))")
    for name, N, Kd, has_res, odt in cases:
        line = []
        for bn, st in cfgs:
            tile = st * 1000 + bn
            graphs = []
            try:
                for s in range(S):
                    x = torch.randn(R, Kd, device=dev).bfloat16(); w = torch.randn(N, Kd, device=dev).bfloat16(); b = torch.randn(N, device=dev)
                    res = torch.randn(R, N, device=dev) if has_res else None
                    outs = [torch.empty(R, N, device=dev, dtype=odt) for _ in range(2)]
                    K.linear(x, w, b, residual=res, out=outs[0], tile_n=tile); torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        for i in range(40): K.linear(x, w, b, residual=res, out=outs[i % 2], tile_n=tile)
                    graphs.append((g, x, w, b, res, outs))
                cur = torch.cuda.current_stream()
                def go():
                    for s in range(S):
                        streams[s].wait_stream(cur)
                        with torch.cuda.stream(streams[s]):
                            graphs[s][0].replay(); graphs[s][0].replay()
                    for s in range(S): cur.wait_stream(streams[s])
                go(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); go(); e1.record(); torch.cuda.synchronize()
                line.append(f"{bn}x{st}:{e0.elapsed_time(e1) * 1e3 / (S * 80):6.2f}")
            except Exception as ex:
                line.append(f"{bn}x{st}:  ERR ")
        print(f"{name:18s} N={N:5d} K={Kd:5d}  " + "  ".join(line), flush=True)
