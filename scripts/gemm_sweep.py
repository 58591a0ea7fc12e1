"""Time the tcgen05 GEMM for the model's shapes under each (tile_n, stages) config, inside a CUDA graph (L2-warm, as
in the decode loop).  Diagnostic for the tile-selection heuristic."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_caption_b200.kernels as K
shapes = [(1536,512,512),(1536,1536,512),(1536,2048,512),(1536,512,2048),(1536,10000,512),
          (18432,512,2048),(18432,1536,512),(18432,2048,512),(18432,512,512),(18432,1024,512),
          (6144,512,512),(6144,1536,512),(6144,2048,512),(6144,512,2048),(4250,512,512),(4250,2048,512),(4250,512,2048),(4250,10000,512),(1800,1536,512),(1800,2048,512)]
cfgs = [(0,0),(64,3),(64,4),(64,6),(128,3),(128,5),(256,3)]
dev = "cuda"
if os.environ.get("SC_PDL") == "0": K.set_pdl(False)
if len(sys.argv) > 1: shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for (M,N,Kd) in shapes:
    x = torch.randn(M,Kd,device=dev).bfloat16(); w = torch.randn(N,Kd,device=dev).bfloat16(); b = torch.randn(N,device=dev)
    outs = [torch.empty(M,N,device=dev,dtype=torch.bfloat16) for _ in range(4)]
    res = []
    for (bn,st) in cfgs:
        tile = st*1000+bn
        try:
            K.linear(x,w,b,out=outs[0],tile_n=tile); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(40): K.linear(x,w,b,out=outs[i%4],tile_n=tile)
            g.replay(); torch.cuda.synchronize()
            e0,e1 = torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1)*1e3/80
            res.append(f"{bn}x{st}:{us:6.1f}us({2*M*N*Kd/us/1e6:5.0f}TF)")
        except Exception as ex:
            res.append(f"{bn}x{st}:ERR")
    print(f"M={M:5d} N={N:5d} K={Kd:4d}  " + "  ".join(res), flush=True)
