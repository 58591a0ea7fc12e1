"""In-graph timing of the weight-gradient GEMM (split-K variants); diagnostic."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_caption_b200.kernels as K
dev = "cuda"
def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (2 * reps)
for (N, Kd, M) in [(512,512,4256),(512,2048,1800),(2048,512,4256),(512,2048,4256),(1536,512,1800),(512,512,1800),(10000,512,4256)]:
    dyT = torch.randn(N, M, device=dev).bfloat16(); xT = torch.randn(Kd, M, device=dev).bfloat16()
    W = torch.randn(N, Kd, device=dev); S = torch.randn(N, Kd, device=dev); dw = torch.zeros(N, Kd, device=dev); ds = torch.zeros(N, Kd, device=dev)
    out = [f"N={N:5d} K={Kd:4d} M={M:4d}"]
    for mode, name in ((2, "bern"), (0, "nomask")):
        for tile in (0, 100128):
            try:
                us = timeit(lambda: K.linear_wgrad(dyT, xT, W, S if mode else None, mode, dw, ds if mode else None, M=M, seed=1, stream_id=3, tile_n=tile))
                out.append(f"{name}/{tile}:{us:6.1f}")
            except Exception as ex:
                out.append(f"{name}/{tile}:ERR")
    for mode, name in ((2, "bern"), (0, "nomask")):
        for cap in (1, 2, 4, 8, 16):
            wsp = torch.empty(cap * N * Kd, device=dev)
            us = timeit(lambda: K.linear_wgrad(dyT, xT, W, S if mode else None, mode, dw, ds if mode else None, M=M, seed=1, stream_id=3, workspace=wsp))
            out.append(f"{name}/ws{cap}:{us:6.1f}")
    print("  ".join(out), flush=True)
