"""In-graph (L2-warm, back-to-back) timing of the decode-loop kernels at R rows; diagnostic only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_caption_b200.kernels as K
from sparse_caption_b200.engine import BeamState
dev = "cuda"
if os.environ.get("SC_PDL") == "0": K.set_pdl(False)
def timeit(name, fn, reps=40):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(f"  {name:34s} {e0.elapsed_time(e1)*1e3/(2*reps):8.2f} us", flush=True)
for B in [int(a) for a in (sys.argv[1:] or ["512", "2048"])]:
    beam, N, d, h, L, V = 3, 36, 512, 8, 16, 10000
    R = B * beam
    print(f"B={B} R={R}")
    bf = dict(device=dev, dtype=torch.bfloat16)
    x = torch.randn(R, d, device=dev); xn = torch.empty(R, d, **bf); a = torch.ones(d, device=dev); b = torch.zeros(d, device=dev)
    timeit("layernorm", lambda: K.layernorm(x, a, b, out=xn))
    qkv = torch.randn(R, 3 * d, **bf); ck = torch.randn(L, R, d, **bf); cv = torch.randn(L, R, d, **bf)
    anc = torch.arange(R, device=dev, dtype=torch.int32).unsqueeze(1).expand(R, L).contiguous(); att = torch.empty(R, d, **bf)
    for t in (0, 8, 15):
        timeit(f"self_attn_step t={t}", lambda: K.self_attn_step(qkv[:, 0:], qkv[:, d:], qkv[:, 2*d:], ck, cv, anc, att, R=R, D=d, h=h,
               n_prev=t, write_slot=t, ldq=3*d, ldk=3*d, ldv=3*d, ldo=d, anc_ld=L, slot_div=1))
    qc = torch.randn(R, d, **bf); mkv = torch.randn(B * N, 2 * d, **bf)
    timeit("cross_attn_step", lambda: K.cross_attn_step(qc, mkv[:, 0:], mkv[:, d:], None, att, B=B, beam=beam, N=N, D=d, h=h, ldq=d, ldm=2*d, ldo=d))
    logits = torch.randn(R, V, device=dev); st = BeamState(B, beam, L, dev); st.reset(2, 0)
    timeit("beam_step t=5", lambda: K.beam_step(logits, st, 5, B=B, beam=beam, V=V, L=L, eos=3, pad=0))
    table = torch.randn(V, d, device=dev); pe = torch.randn(L + 2, d, device=dev); tok = torch.randint(0, V, (R,), device=dev, dtype=torch.int32)
    timeit("embed_pe", lambda: K.embed_pe(tok, table, pe, T=1, pos0=3, out=x))
    for (Nn, Kd) in [(512, 512), (1536, 512), (2048, 512), (512, 2048), (10000, 512)]:
        xx = torch.randn(R, Kd, **bf); w = torch.randn(Nn, Kd, **bf); bb = torch.randn(Nn, device=dev)
        y = torch.empty(R, Nn, **bf)
        timeit(f"gemm {R}x{Nn}x{Kd} auto", lambda: K.linear(xx, w, bb, out=y))
