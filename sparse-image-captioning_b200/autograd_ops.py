"""Layer-level autograd bindings of the masked kernels: what MaskedLinear / MaskedEmbedding call.

Forward: sc_linear / sc_embed_pe with the mask applied in the operand prologue / gather.
Backward (K2): dX through (W (.) m)^T, dW and dS from the fused weight-gradient epilogue with the SAME mask
regenerated from (seed, stream) — the straight-through estimators of sparse_caption/pruning/sampler.py:10-66.
"""
import torch

from . import kernels as K
from . import sampler


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: the B200 path runs on CUDA tensors only (there is no CPU fallback)")


class _MaskedLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, S, bias, mode, bypass, precision, seed, stream):
        Kd = W.shape[1]
        x2 = x.reshape(-1, Kd)
        adt = torch.bfloat16 if precision == "bf16" else torch.float32
        if adt == torch.bfloat16 and Kd % 8 != 0:
            adt = torch.float32  # TMA rows need 16-byte strides; tiny layers (e.g. the 64->1 WG heads) use the fp32 kernel
        xa = x2.contiguous()
        if xa.dtype != adt:
            xa = K.cast_bf16(xa.float().contiguous()) if adt == torch.bfloat16 else xa.float()
        y = K.linear(xa, W.detach(), None if bias is None else bias.detach(), mask=S.detach(), mask_mode=mode, seed=seed, stream_id=stream)
        ctx.save_for_backward(xa, W, S)
        ctx.meta = (mode, bypass, seed, stream, bias is not None, x.shape)
        return y.reshape(x.shape[:-1] + (W.shape[0],))

    @staticmethod
    def backward(ctx, dy):
        xa, W, S = ctx.saved_tensors
        mode, bypass, seed, stream, has_bias, xshape = ctx.meta
        N, Kd = W.shape
        M = xa.shape[0]
        Mp = K.pad8(M)
        adt = xa.dtype
        g = dy.reshape(M, N).float().contiguous()
        gb = torch.empty(M, N, device=g.device, dtype=adt)
        gT = torch.zeros(N, Mp, device=g.device, dtype=adt)
        K.prep_grad(g, out=gb, outT=gT)
        db = K.colsum(gb, torch.empty(N, device=g.device)) if has_bias else None
        dx = None
        if ctx.needs_input_grad[0]:
            if adt == torch.bfloat16 and N % 8 != 0:
                raise RuntimeError("masked_linear backward (bf16): out_features must be a multiple of 8")
            wT = torch.empty(Kd, N, device=g.device, dtype=adt)
            K.apply_mask_transposed(W.detach(), S.detach(), mode, wT, seed=seed, stream_id=stream)
            dx = K.linear(gb, wT).reshape(xshape)
        xT = torch.zeros(Kd, Mp, device=g.device, dtype=adt)
        K.transpose(xa, xT)
        dW = torch.empty_like(W)
        dS = torch.empty_like(S) if ctx.needs_input_grad[2] else None
        K.linear_wgrad(gT, xT, W.detach(), S.detach(), mode, dW, dS, M=Mp, seed=seed, stream_id=stream, bypass=bypass)
        return dx, dW, dS, db, None, None, None, None, None


def masked_linear(x, W, S, bias, mode, bypass, precision):
    _need_cuda(x, "MaskedLinear")
    seed, stream = sampler.next_mask_stream() if mode == K.MASK_BERNOULLI else (0, 0)
    return _MaskedLinearFn.apply(x, W, S, bias, mode, bypass, precision, seed, stream)


class _MaskedEmbeddingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, W, S, mode, bypass, seed, stream):
        V, D = W.shape
        tok = ids.reshape(-1).to(torch.int32).contiguous()
        zero_pe = torch.zeros(1, D, device=W.device)
        out = K.embed_pe(tok, W.detach(), zero_pe, T=1, pos0=0, mask=S.detach(), mask_mode=mode, seed=seed, stream_id=stream)
        out = out / (D ** 0.5)  # sc_embed_pe folds the sqrt(d) of InputEmbedding; the bare layer does not scale
        ctx.save_for_backward(tok, W, S)
        ctx.meta = (mode, bypass, seed, stream)
        return out.reshape(ids.shape + (D,))

    @staticmethod
    def backward(ctx, dy):
        tok, W, S = ctx.saved_tensors
        mode, bypass, seed, stream = ctx.meta
        V, D = W.shape
        dtab = torch.zeros(V, D, device=W.device)
        K.embedding_bwd(tok, dy.reshape(-1, D).float().contiguous(), dtab, 1.0)
        dW = torch.empty_like(W)
        dS = torch.empty_like(S) if ctx.needs_input_grad[2] else None
        K.mask_grad(dtab, W.detach(), S.detach(), mode, dW, dS, seed=seed, stream_id=stream, bypass=bypass)
        return None, dW, dS, None, None, None, None


def masked_embedding(ids, W, S, mode, bypass):
    _need_cuda(W, "MaskedEmbedding")
    seed, stream = sampler.next_mask_stream() if mode == K.MASK_BERNOULLI else (0, 0)
    return _MaskedEmbeddingFn.apply(ids.to(W.device), W, S, mode, bypass, seed, stream)
