"""Wall time per launch of each decode-loop kernel when S streams run it concurrently (pipeline slots): elapsed / (S*80)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_caption_b200.kernels as K
from sparse_caption_b200.engine import BeamState
dev = "cuda"
B, beam, N, d, h, L, V = 512, 3, 36, 512, 8, 16, 10000
R = B * beam
bf = dict(device=dev, dtype=torch.bfloat16)
def make(kind):
    if kind == "layernorm":
        x = torch.randn(R, d, device=dev); xn = torch.empty(R, d, **bf); a = torch.ones(d, device=dev); b = torch.zeros(d, device=dev)
        return lambda i: K.layernorm(x, a, b, out=xn)
    if kind.startswith("self_attn"):
        t = int(kind.split("=")[1])
        qkv = torch.randn(R, 3 * d, **bf); ck = torch.randn(L, R, d, **bf); cv = torch.randn(L, R, d, **bf)
        anc = torch.arange(R, device=dev, dtype=torch.int32).unsqueeze(1).expand(R, L).contiguous(); att = torch.empty(R, d, **bf)
        return lambda i: K.self_attn_step(qkv[:, 0:], qkv[:, d:], qkv[:, 2*d:], ck, cv, anc, att, R=R, D=d, h=h, n_prev=t, write_slot=t,
                                          ldq=3*d, ldk=3*d, ldv=3*d, ldo=d, anc_ld=L, slot_div=1)
    if kind == "cross_attn":
        qc = torch.randn(R, d, **bf); mkv = torch.randn(B * N, 2 * d, **bf); att = torch.empty(R, d, **bf)
        return lambda i: K.cross_attn_step(qc, mkv[:, 0:], mkv[:, d:], None, att, B=B, beam=beam, N=N, D=d, h=h, ldq=d, ldm=2*d, ldo=d)
    if kind == "beam_step":
        logits = torch.randn(R, V, device=dev); st = BeamState(B, beam, L, dev); st.reset(2, 0)
        return lambda i: K.beam_step(logits, st, 5, B=B, beam=beam, V=V, L=L, eos=3, pad=0)
    if kind == "embed_pe":
        table = torch.randn(V, d, device=dev); pe = torch.randn(L + 2, d, device=dev); tok = torch.randint(0, V, (R,), device=dev, dtype=torch.int32)
        x = torch.empty(R, d, device=dev)
        return lambda i: K.embed_pe(tok, table, pe, T=1, pos0=3, out=x)
    if kind == "generator":
        x = torch.randn(R, d, **bf); w = torch.randn(V, d, **bf); b = torch.randn(V, device=dev); y = torch.empty(R, V, device=dev)
        return lambda i: K.linear(x, w, b, out=y)
kinds = ["layernorm", "self_attn t=0", "self_attn t=8", "self_attn t=15", "cross_attn", "beam_step", "embed_pe", "generator"]
for S in (1, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(S)]
    print(f"--- {S} concurrent streams: us of wall time per launch")
    for kind in kinds:
        graphs = []
        for s in range(S):
            fn = make(kind); fn(0); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(40): fn(i)
            graphs.append((g, fn))
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
        print(f"  {kind:16s} {e0.elapsed_time(e1) * 1e3 / (S * 80):7.2f}", flush=True)
