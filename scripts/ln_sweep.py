"""In-graph timing of plain vs LayerNorm-folded GEMM variants (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_caption_b200.kernels as K
dev = "cuda"
if os.environ.get("SC_PDL") == "0": K.set_pdl(False)
def timeit(fn, reps=40):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (2 * reps)
shapes = [(1536,512,512),(1536,1536,512),(1536,2048,512),(1536,512,2048),(1536,10000,512),(18432,1536,512),(18432,512,2048),(18432,2048,512),(18432,512,512)]
for (M,N,Kd) in shapes:
    x = torch.randn(M,Kd,device=dev).bfloat16(); w = torch.randn(N,Kd,device=dev).bfloat16(); b = torch.randn(N,device=dev)
    res = torch.randn(M,N,device=dev); y32 = torch.empty(M,N,device=dev); yb = torch.empty(M,N,device=dev,dtype=torch.bfloat16)
    st_in = torch.rand(M,Kd//32,2,device=dev)+1; lnc = torch.randn(N,device=dev)
    out = [f"M={M:5d} N={N:5d} K={Kd:4d}"]
    out.append(f"plain->bf16 {timeit(lambda: K.linear(x,w,b,out=yb)):6.1f}")
    out.append(f"plain+res->f32 {timeit(lambda: K.linear(x,w,b,residual=y32,out=y32)):6.1f}")
    out.append(f"ln->bf16 {timeit(lambda: K.linear_ln(x,w,b,out=yb,ln_stats=st_in,ln_c=lnc)):6.1f}")
    if N % 32 == 0:
        st_out = torch.empty(M,N//32,2,device=dev)
        out.append(f"produce {timeit(lambda: K.linear_ln(x,w,b,residual=y32,out=y32,out_bf16=yb,stats_out=st_out)):6.1f}")
    xx = torch.randn(M,512,device=dev); a = torch.ones(512,device=dev); xn = torch.empty(M,512,device=dev,dtype=torch.bfloat16)
    out.append(f"LN512 {timeit(lambda: K.layernorm(xx,a,a,out=xn)):6.1f}")
    print("  ".join(out), flush=True)
