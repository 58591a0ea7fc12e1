"""Tensor-level wrappers over the C ABI (one Python function per entry point of include/sc_b200.h).

All tensors must be CUDA tensors; outputs are allocated by the caller or here with torch (plumbing only).
Every function launches on torch's current stream, so calls are CUDA-graph capturable.
"""
import math

import torch

from . import lib
from .lib import F32, BF16, MASK_NONE, MASK_ROUND, MASK_BERNOULLI, MASK_RAW, MASK_UNIFORM  # noqa: F401


def _chk(t, name):
    if t is not None:
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor: the B200 path has no CPU fallback")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be contiguous")


def linear(x, w, bias=None, *, mask=None, mask_mode=MASK_NONE, uniforms=None, seed=0, stream_id=0, residual=None,
           relu=False, out=None, out_dtype=None, tile_n=0):
    """y = epilogue(x @ (w (.) mask)^T): sc_linear.  x [M,K] bf16|fp32, w [N,K] bf16|fp32."""
    M, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K, (w.shape, x.shape)
    for t, n in ((x, "x"), (w, "w"), (mask, "mask"), (uniforms, "uniforms"), (bias, "bias"), (residual, "residual")):
        _chk(t, n)
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=out_dtype or torch.float32)
    if mask is None:
        mask_mode = MASK_NONE
    lib.call("sc_linear", lib.ptr(x), lib.dtype_code(x.dtype), lib.ptr(w), lib.dtype_code(w.dtype), lib.ptr(mask),
             mask_mode, lib.ptr(uniforms), seed, stream_id, lib.ptr(bias), lib.ptr(residual), lib.ptr(out),
             lib.dtype_code(out.dtype), M, N, K, int(relu), tile_n, lib.stream(),
             meta=("gemm_bf16" if x.dtype == torch.bfloat16 else "gemm_f32", M, N, K, x.element_size(), w.element_size(),
                   out.element_size(), mask is not None))
    return out


class CsrWeight:
    """CSR form of a pruned [N,K] weight: row_ptr int32 [N+1], col uint16 [nnz] (stored in an int16 tensor),
    values in the activation dtype.  Built from a dense (already masked) tensor; also accepts the reference's
    COO ``to_sparse()`` tensors (pruning/prune.py:200-221)."""

    def __init__(self, w, dtype):
        if w.is_sparse:
            w = w.to_dense()
        w = w.detach()
        N, K = w.shape
        assert K <= 65536
        nz = w != 0
        counts = nz.sum(1)
        row_ptr = torch.zeros(N + 1, dtype=torch.int64, device=w.device)
        row_ptr[1:] = torch.cumsum(counts, 0)
        idx = nz.nonzero(as_tuple=False)  # row-major order -> already grouped by row, columns ascending
        self.row_ptr = row_ptr.to(torch.int32).contiguous()
        self.col = idx[:, 1].to(torch.int32).to(torch.int16).contiguous()  # bit pattern of uint16
        self.val = w[nz].to(dtype).contiguous()
        self.shape = (N, K)
        self.nnz = int(self.val.numel())
        if self.nnz == 0:  # keep pointers valid
            self.col = torch.zeros(1, dtype=torch.int16, device=w.device)
            self.val = torch.zeros(1, dtype=dtype, device=w.device)


def csr_spmm(x, csr, bias=None, *, residual=None, relu=False, out=None, out_dtype=None):
    M, K = x.shape
    N = csr.shape[0]
    assert csr.shape[1] == K and csr.val.dtype == x.dtype
    _chk(x, "x")
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=out_dtype or torch.float32)
    lib.call("sc_csr_spmm", lib.ptr(x), lib.dtype_code(x.dtype), lib.ptr(csr.row_ptr), lib.ptr(csr.col), lib.ptr(csr.val),
             lib.ptr(bias), lib.ptr(residual), lib.ptr(out), lib.dtype_code(out.dtype), M, N, K, int(relu), lib.stream(),
             meta=("csr_spmm", M, N, K, x.element_size(), csr.nnz, out.element_size()))
    return out


def layernorm(x, a, b, *, eps=1e-6, out=None, out_dtype=None):
    rows, D = x.shape
    _chk(x, "x")
    assert x.dtype == torch.float32
    if out is None:
        out = torch.empty(rows, D, device=x.device, dtype=out_dtype or torch.float32)
    lib.call("sc_layernorm", lib.ptr(x), lib.ptr(a), lib.ptr(b), lib.ptr(out), lib.dtype_code(out.dtype), rows, D, eps,
             lib.stream())
    return out


def embed_pe(tokens, table, pe, *, T=1, pos0=0, mask=None, mask_mode=MASK_NONE, uniforms=None, seed=0, stream_id=0,
             out=None, out_dtype=None):
    rows = tokens.numel()
    V, D = table.shape
    assert tokens.dtype == torch.int32
    if out is None:
        out = torch.empty(rows, D, device=table.device, dtype=out_dtype or torch.float32)
    if mask is None:
        mask_mode = MASK_NONE
    lib.call("sc_embed_pe", lib.ptr(tokens), lib.ptr(table), lib.ptr(mask), mask_mode, lib.ptr(uniforms), seed, stream_id,
             lib.ptr(pe), lib.ptr(out), lib.dtype_code(out.dtype), rows, D, V, T, pos0, math.sqrt(D), lib.stream())
    return out


def apply_mask(w, mask, mask_mode, *, uniforms=None, seed=0, stream_id=0, out_dtype=torch.float32):
    _chk(w, "w")
    out = torch.empty(w.shape, device=w.device, dtype=out_dtype)
    if mask is None:
        mask_mode = MASK_NONE
    lib.call("sc_apply_mask", lib.ptr(w), lib.ptr(mask), mask_mode, lib.ptr(uniforms), seed, stream_id, lib.ptr(out),
             lib.dtype_code(out_dtype), w.numel(), lib.stream())
    return out


def mask_count(logits_list):
    """sum over tensors of sum(rint(sigmoid(S))) -> int64 CUDA scalar tensor."""
    dev = logits_list[0].device
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    for s in logits_list:
        _chk(s, "logits")
        lib.call("sc_mask_count", lib.ptr(s), s.numel(), lib.ptr(cnt), lib.stream())
    return cnt


def cast_bf16(x, out=None):
    _chk(x, "x")
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    lib.call("sc_cast_f32_bf16", lib.ptr(x), lib.ptr(out), x.numel(), lib.stream())
    return out


def mask_rows(x, row_mask):
    rows, D = x.shape
    lib.call("sc_mask_rows", lib.ptr(x), lib.ptr(row_mask), rows, D, lib.stream())
    return x


def box_attention(q, k, v, boxes, wg_w, wg_b, att_mask, out, *, B, N, h, dk, ldq, ldk, ldv, ldo, trig=True,
                  wave_len=1000.0):
    lib.call("sc_box_attention_fwd", lib.ptr(q), lib.ptr(k), lib.ptr(v), ldq, ldk, ldv, lib.dtype_code(out.dtype),
             lib.ptr(boxes), lib.ptr(wg_w), lib.ptr(wg_b), lib.ptr(att_mask), lib.ptr(out), ldo, B, N, h, dk, int(trig),
             wave_len, lib.stream())
    return out


def self_attn_step(q, k, v, cache_k, cache_v, anc, out, *, R, D, h, n_prev, write_slot, ldq, ldk, ldv, ldo, anc_ld,
                   slot_div=1):
    lib.call("sc_decode_self_attn_step", lib.ptr(q), lib.ptr(k), lib.ptr(v), ldq, ldk, ldv, lib.dtype_code(out.dtype),
             lib.ptr(cache_k), lib.ptr(cache_v), lib.ptr(anc), anc_ld, slot_div, lib.ptr(out), ldo, R, D, h, n_prev,
             write_slot, lib.stream())
    return out


def cross_attn_step(q, mem_k, mem_v, att_mask, out, *, B, beam, N, D, h, ldq, ldm, ldo):
    lib.call("sc_decode_cross_attn_step", lib.ptr(q), ldq, lib.ptr(mem_k), lib.ptr(mem_v), ldm, lib.dtype_code(out.dtype),
             lib.ptr(att_mask), lib.ptr(out), ldo, B, beam, N, D, h, lib.stream())
    return out


def beam_step(logits, st, t, *, B, beam, V, L, eos, pad, temperature=1.0, constraint=0, penalty_kind=0, penalty_alpha=0.0):
    """``st``: BeamState.  Reads buffers ``t % 2`` and writes ``(t+1) % 2``."""
    i, o = t & 1, (t + 1) & 1
    lib.call("sc_beam_step", lib.ptr(logits), B, beam, V, L, t, eos, pad, float(temperature), int(constraint),
             int(penalty_kind), float(penalty_alpha), lib.ptr(st.seq[i]), lib.ptr(st.seq[o]), lib.ptr(st.lp[i]),
             lib.ptr(st.lp[o]), lib.ptr(st.sum), lib.ptr(st.anc[i]), lib.ptr(st.anc[o]), lib.ptr(st.tokens),
             lib.ptr(st.done_seq), lib.ptr(st.done_lp), lib.ptr(st.done_p), lib.ptr(st.done_count), lib.stream())


def greedy_step(logits, st, t, *, R, V, L, eos, constraint=0):
    lib.call("sc_greedy_step", lib.ptr(logits), R, V, L, t, eos, int(constraint), lib.ptr(st.seq), lib.ptr(st.lp),
             lib.ptr(st.tokens), lib.ptr(st.unfinished), lib.ptr(st.live), lib.stream())


def cache_reorder(src, idx, out=None):
    """dst[r] = src[idx[r]] along dim 0 (K8)."""
    rows = idx.numel()
    row_bytes = src[0].numel() * src.element_size()
    if out is None:
        out = torch.empty((rows,) + tuple(src.shape[1:]), device=src.device, dtype=src.dtype)
    assert idx.dtype == torch.int32
    lib.call("sc_cache_reorder", lib.ptr(src), lib.ptr(out), lib.ptr(idx), rows, row_bytes, lib.stream())
    return out
