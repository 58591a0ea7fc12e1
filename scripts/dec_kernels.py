"""Every kernel of one decode layer-step at the coalesced device batch (R = images x G x beam rows), timed alone inside a CUDA
graph (activations L2-warm as in the real chain; K/V caches rotate over buffers larger than L2), next to its HBM / tensor floor.

    python scripts/dec_kernels.py [--images 512] [--coalesce 5] [--beam 3] [--hints 20003256,3256,3128,0]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from bench import _time_graph, load_peaks  # noqa: E402
from sparse_caption_b200 import kernels as K, lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=512)
    ap.add_argument("--coalesce", type=int, default=5)
    ap.add_argument("--beam", type=int, default=3)
    ap.add_argument("--hints", default="20003256,3256,3128,5128,0")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    lib.load()
    dev = torch.device("cuda", 0)
    peaks = load_peaks()
    B = a.images * a.coalesce
    R, d, ff, h, N, L, V = B * a.beam, 512, 2048, 8, 36, 16, 10000
    bf = dict(device=dev, dtype=torch.bfloat16)
    rows = []

    def rec(name, us, nbytes=0, flop=0):
        r = {"kernel": name, "us": round(us, 2)}
        if nbytes:
            r["GB/s"] = round(nbytes / us / 1e3)
            r["hbm_floor_us"] = round(nbytes / peaks["hbm"] / 1e3, 2)
        if flop:
            r["TF/s"] = round(flop / us / 1e6)
            r["tensor_floor_us"] = round(flop / peaks["tf_sus"] / 1e6, 2)
        rows.append(r)
        print(json.dumps(r), flush=True)

    x32 = torch.randn(R, d, device=dev)
    x32b = torch.randn(R, d, device=dev)
    xn = torch.randn(R, d, **bf)
    a2, b2 = torch.ones(d, device=dev), torch.zeros(d, device=dev)
    rec("layernorm", _time_graph(lambda i: K.layernorm(x32, a2, b2, out=xn), dev), R * d * 6)
    hints = [int(s) for s in a.hints.split(",")]
    shapes = [("qkv", 3 * d, d, 2, False, False), ("o", d, d, 4, True, False), ("cq", d, d, 2, False, False),
              ("ff1", ff, d, 2, False, True), ("ff2", d, ff, 4, True, False)]
    for name, Nn, Kd, ys, has_res, relu in shapes:
        xin = torch.randn(R, Kd, **bf)
        w = torch.randn(Nn, Kd, **bf)
        bias = torch.randn(Nn, device=dev)
        out = torch.empty(R, Nn, device=dev, dtype=torch.bfloat16 if ys == 2 else torch.float32)
        res = x32 if has_res else None
        outp = x32 if has_res else out  # (in place on the residual stream, as the engine runs it)
        nbytes = R * Kd * 2 + Nn * Kd * 2 + R * Nn * ys + (R * Nn * 4 if has_res else 0)
        for hint in hints:
            try:
                us = _time_graph(lambda i: K.linear(xin, w, bias, residual=res, relu=relu, out=outp, tile_n=hint), dev)
                rec(f"gemm {name} {R}x{Nn}x{Kd} hint={hint}", us, nbytes, 2.0 * R * Nn * Kd)
            except Exception as ex:  # a hint that has no instantiation
                print(f"# {name} hint {hint}: {ex}", flush=True)
    wg = torch.randn(V, d, **bf)
    bg = torch.randn(V, device=dev)
    part = torch.empty(R, K.linear_topk_parts(V), 12, device=dev)
    rec(f"generator topk {R}x{V}x{d}", _time_graph(lambda i: K.linear_topk(xn, wg, bg, part, candidates=a.beam), dev), R * d * 2 + V * d * 2,
        2.0 * R * V * d)
    C = 3
    qkv = torch.randn(R, 3 * d, **bf)
    ck = [torch.randn(L, R, d, **bf) for _ in range(C)]
    cv = [torch.randn(L, R, d, **bf) for _ in range(C)]
    anc = torch.arange(R, device=dev, dtype=torch.int32).unsqueeze(1).expand(R, L).contiguous()
    att = torch.empty(R, d, **bf)
    for t in (0, L // 2, L - 1):
        rec(f"self_attn_step t={t}",
            _time_graph(lambda i, t=t: K.self_attn_step(qkv[:, 0:], qkv[:, d:], qkv[:, 2 * d:], ck[i % C], cv[i % C], anc, att, R=R, D=d, h=h,
                                                        n_prev=t, write_slot=t, ldq=3 * d, ldk=3 * d, ldv=3 * d, ldo=d, anc_ld=L, slot_div=1), dev),
            R * (t + 1) * 2 * d * 2 + R * 3 * d * 2 + 2 * R * d * 2 + R * d * 2)
    mkv = [torch.randn(B * N, 2 * d, **bf) for _ in range(C)]
    qc = torch.randn(R, d, **bf)
    rec("cross_attn_step", _time_graph(lambda i: K.cross_attn_step(qc, mkv[i % C][:, 0:], mkv[i % C][:, d:], None, att, B=B, beam=a.beam, N=N,
                                                                    D=d, h=h, ldq=d, ldm=2 * d, ldo=d), dev),
        B * N * 2 * d * 2 + 2 * R * d * 2)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
