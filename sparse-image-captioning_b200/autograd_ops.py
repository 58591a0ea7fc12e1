"""Layer-level autograd bindings of the kernels: what the module tree (MaskedLinear / MaskedEmbedding / LayerNorm /
attention modules / generator) calls, so that ``model(**data)`` -> ``loss.backward()`` -> ``clip_gradient`` ->
``optimizer.step()`` of the reference loop (scripts/train_n_prune_transformer.py:136-153) runs unmodified.

Every forward and every backward is one or a few launches through include/sc_b200.h; tensors cross module boundaries as
fp32 (like the reference) and are cast to the GEMM operand dtype inside the functions ("bf16" precision) or kept ("fp32").
Masked layers: the mask is applied in the operand prologue (sc_linear) / embedding gather; backward (K2): dX through
(W (.) m)^T, dW and dS from the fused weight-gradient epilogue with the SAME mask regenerated from (seed, stream) - the
straight-through estimators of sparse_caption/pruning/sampler.py:10-66.

This is the general (autograd, any sharing pattern) path.  The fused training engine (trainer.OrtTrainer) and the
inference engine (engine.OrtEngine) are the fast paths over the same kernels.
"""
import math

import torch

from . import kernels as K
from . import sampler


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: the B200 path runs on CUDA tensors only (there is no CPU fallback)")


def _adt(precision):
    return torch.bfloat16 if precision == "bf16" else torch.float32


def _operand(x2, adt):
    """fp32 / bf16 2-D tensor -> contiguous GEMM operand in ``adt``."""
    xa = x2.contiguous()
    if xa.dtype != adt:
        xa = K.cast_bf16(xa.float().contiguous()) if adt == torch.bfloat16 else xa.float()
    return xa


class _MaskedLinearFn(torch.autograd.Function):
    """y = act(x (W (.) m)^T + b).  S may be None (plain nn.Linear)."""

    @staticmethod
    def forward(ctx, x, W, S, bias, mode, bypass, precision, seed, stream, relu, uniforms):
        Kd = W.shape[1]
        x2 = x.reshape(-1, Kd)
        adt = _adt(precision)
        if adt == torch.bfloat16 and Kd % 8 != 0:
            adt = torch.float32  # TMA rows need 16-byte strides; tiny layers (e.g. the 64->1 WG heads) use the fp32 kernel
        xa = _operand(x2, adt)
        Sd = None if S is None else S.detach()
        y = K.linear(xa, W.detach(), None if bias is None else bias.detach(), mask=Sd, mask_mode=mode if S is not None else K.MASK_NONE,
                     uniforms=uniforms, seed=seed, stream_id=stream, relu=relu)
        ctx.save_for_backward(xa, W, S, y if relu else None, uniforms)
        ctx.meta = (mode if S is not None else K.MASK_NONE, bypass, seed, stream, bias is not None, x.shape, relu)
        return y.reshape(x.shape[:-1] + (W.shape[0],))

    @staticmethod
    def backward(ctx, dy):
        xa, W, S, y, U = ctx.saved_tensors
        mode, bypass, seed, stream, has_bias, xshape, relu = ctx.meta
        N, Kd = W.shape
        M = xa.shape[0]
        Mp = K.pad8(M)
        adt = xa.dtype
        Sd = None if S is None else S.detach()
        g = dy.reshape(M, N).float().contiguous()
        gb = torch.empty(M, N, device=g.device, dtype=adt)
        gT = torch.zeros(N, Mp, device=g.device, dtype=adt)
        K.prep_grad(g, h=y if relu else None, out=gb, outT=gT)  # (ReLU: keep where the saved output is non-zero)
        db = K.colsum(gb, torch.empty(N, device=g.device)) if has_bias else None
        dx = None
        if ctx.needs_input_grad[0]:
            if adt == torch.bfloat16 and N % 8 != 0:
                raise RuntimeError("masked_linear backward (bf16): out_features must be a multiple of 8")
            wT = torch.empty(Kd, N, device=g.device, dtype=adt)
            K.apply_mask_transposed(W.detach(), Sd, mode, wT, uniforms=U, seed=seed, stream_id=stream)
            dx = K.linear(gb, wT).reshape(xshape)
        xT = torch.zeros(Kd, Mp, device=g.device, dtype=adt)
        K.transpose(xa, xT)
        dW = torch.empty_like(W)
        dS = torch.empty_like(S) if (S is not None and ctx.needs_input_grad[2]) else None
        K.linear_wgrad(gT, xT, W.detach(), Sd, mode, dW, dS, M=Mp, uniforms=U, seed=seed, stream_id=stream, bypass=bypass)
        return dx, dW, dS, db, None, None, None, None, None, None, None


def _mask_draw(mode, S):
    """(mode, seed, stream, uniforms) of one forward call: a fresh Philox stream per sampled tensor per call (the reference
    re-samples on every MaskedLinear.forward, masked_layer.py:92-102), or injected uniforms (parity tests)."""
    if mode == K.MASK_BERNOULLI:
        u = sampler.injected_uniforms(S)
        if u is not None:
            return K.MASK_UNIFORM, 0, 0, u
        seed, stream = sampler.next_mask_stream()
        return mode, seed, stream, None
    return mode, 0, 0, None


def masked_linear(x, W, S, bias, mode, bypass, precision, relu=False):
    _need_cuda(x, "MaskedLinear")
    mode, seed, stream, u = _mask_draw(mode, S)
    return _MaskedLinearFn.apply(x, W, S, bias, mode, bypass, precision, seed, stream, relu, u)


def linear(x, W, bias, precision, relu=False):
    """Plain nn.Linear through the same kernels (dense `relation_transformer` class)."""
    _need_cuda(x, "Linear")
    return _MaskedLinearFn.apply(x, W, None, bias, K.MASK_NONE, False, precision, 0, 0, relu, None)


class _MaskedWeightFn(torch.autograd.Function):
    """W (.) m materialised (small tensors only: the 8 x [1, 64] geometry heads of an encoder layer)."""

    @staticmethod
    def forward(ctx, W, S, mode, bypass, seed, stream, uniforms):
        ctx.save_for_backward(W, S, uniforms)
        ctx.meta = (mode, bypass, seed, stream)
        return K.apply_mask(W.detach().contiguous(), S.detach().contiguous(), mode, uniforms=uniforms, seed=seed, stream_id=stream)

    @staticmethod
    def backward(ctx, g):
        W, S, U = ctx.saved_tensors
        mode, bypass, seed, stream = ctx.meta
        dW = torch.empty_like(W)
        dS = torch.empty_like(S) if ctx.needs_input_grad[1] else None
        K.mask_grad(g.float().contiguous(), W.detach().contiguous(), S.detach().contiguous(), mode, dW, dS, uniforms=U, seed=seed,
                    stream_id=stream, bypass=bypass)
        return dW, dS, None, None, None, None, None


def masked_weight(W, S, mode, bypass):
    _need_cuda(W, "masked weight")
    mode, seed, stream, u = _mask_draw(mode, S)
    return _MaskedWeightFn.apply(W, S, mode, bypass, seed, stream, u)


class _MaskedEmbeddingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, W, S, mode, bypass, seed, stream, uniforms):
        V, D = W.shape
        tok = ids.reshape(-1).to(torch.int32).contiguous()
        zero_pe = torch.zeros(1, D, device=W.device)
        Sd = None if S is None else S.detach()
        out = K.embed_pe(tok, W.detach(), zero_pe, T=1, pos0=0, mask=Sd, mask_mode=mode if S is not None else K.MASK_NONE,
                         uniforms=uniforms, seed=seed, stream_id=stream)
        out = out / (D ** 0.5)  # sc_embed_pe folds the sqrt(d) of InputEmbedding; the bare layer does not scale
        ctx.save_for_backward(tok, W, S, uniforms)
        ctx.meta = (mode if S is not None else K.MASK_NONE, bypass, seed, stream)
        return out.reshape(ids.shape + (D,))

    @staticmethod
    def backward(ctx, dy):
        tok, W, S, U = ctx.saved_tensors
        mode, bypass, seed, stream = ctx.meta
        V, D = W.shape
        dtab = torch.zeros(V, D, device=W.device)
        K.embedding_bwd(tok, dy.reshape(-1, D).float().contiguous(), dtab, 1.0)
        if S is None:
            return None, dtab, None, None, None, None, None, None
        dW = torch.empty_like(W)
        dS = torch.empty_like(S) if ctx.needs_input_grad[2] else None
        K.mask_grad(dtab, W.detach(), S.detach(), mode, dW, dS, uniforms=U, seed=seed, stream_id=stream, bypass=bypass)
        return None, dW, dS, None, None, None, None, None


def masked_embedding(ids, W, S, mode, bypass):
    _need_cuda(W, "MaskedEmbedding")
    mode, seed, stream, u = _mask_draw(mode, S)
    return _MaskedEmbeddingFn.apply(ids.to(W.device), W, S, mode, bypass, seed, stream, u)


def embedding(ids, W):
    _need_cuda(W, "Embedding")
    return _MaskedEmbeddingFn.apply(ids.to(W.device), W, None, K.MASK_NONE, False, 0, 0, None)


# ---- LayerNorm (transformer.py:329-341: unbiased std, eps added to the std) -----------------------------------------
class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, a, b, eps):
        D = x.shape[-1]
        x2 = x.reshape(-1, D).float().contiguous()
        y = K.layernorm(x2, a.detach().float().contiguous(), b.detach().float().contiguous(), eps=eps)
        ctx.save_for_backward(x2, a)
        ctx.eps, ctx.shape = eps, x.shape
        return y.reshape(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, a = ctx.saved_tensors
        D = x2.shape[1]
        g = dy.reshape(-1, D).float().contiguous()
        dx = torch.empty_like(x2)
        da, db = torch.zeros(D, device=x2.device), torch.zeros(D, device=x2.device)
        K.layernorm_bwd(x2, a.detach().float().contiguous(), g, dx, da, db, eps=ctx.eps)
        return dx.reshape(ctx.shape), da, db, None


def layer_norm(x, a, b, eps=1e-6):
    _need_cuda(x, "LayerNorm")
    return _LayerNormFn.apply(x, a, b, eps)


# ---- dropout (Philox, regenerated in the backward) -------------------------------------------------------------------
class _DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, seed, stream):
        x2 = x.reshape(-1, x.shape[-1]).float().contiguous()
        y = torch.empty_like(x2)
        K.prep_grad(x2, out=y, p=p, seed=seed, stream_id=stream)
        ctx.meta = (p, seed, stream, x.shape)
        return y.reshape(x.shape)

    @staticmethod
    def backward(ctx, dy):
        p, seed, stream, shape = ctx.meta
        g = dy.reshape(-1, shape[-1]).float().contiguous()
        dx = torch.empty_like(g)
        K.prep_grad(g, out=dx, p=p, seed=seed, stream_id=stream)
        return dx.reshape(shape), None, None, None


def dropout(x, p, training):
    """nn.Dropout with the framework's Philox stream (inverted dropout, scale 1 / (1 - p))."""
    if not training or p <= 0.0:
        return x
    _need_cuda(x, "Dropout")
    seed, stream = sampler.next_dropout_stream()
    return _DropoutFn.apply(x, float(p), seed, stream)


# ---- attention with saved probabilities (transformer.py:285-295, relation_transformer.py:258-293) ----------------------
class _AttentionFn(torch.autograd.Function):
    """q [G*Tq, h*dk], k / v [G*Tk, h*dk] (2-D, row = (group, position), head-major columns) -> out [G*Tq, h*dk].
    ``key_valid`` fp32 [G, Tk] (0 = masked key) or None; ``bias`` fp32 [G, h, Tq, Tk] additive (log geometry weights) or
    None; ``causal_T`` > 0: keys after the query's position are masked."""

    @staticmethod
    def forward(ctx, q, k, v, bias, key_valid, G, Tq, Tk, h, causal_T, p, seed, stream, precision):
        d = q.shape[1]
        dk = d // h
        adt = _adt(precision)
        qa, ka, va = _operand(q, adt), _operand(k, adt), _operand(v, adt)
        out = torch.empty(G * Tq, d, device=q.device, dtype=adt)
        probs = torch.empty(G, h, Tq, Tk, device=q.device)
        bias_c = None if bias is None else bias.detach().float().contiguous()
        kv = None if key_valid is None else key_valid.float().contiguous()
        K.attention_fwd(qa, ka, va, out, probs, G=G, Tq=Tq, Tk=Tk, h=h, dk=dk, ldq=d, ldk=d, ldv=d, ldo=d, key_valid=kv, bias=bias_c,
                        causal_T=causal_T, p=p, seed=seed, stream_id=stream)
        ctx.save_for_backward(qa, ka, va, probs)
        ctx.meta = (G, Tq, Tk, h, dk, p, seed, stream, adt, bias is not None)
        ctx.mark_non_differentiable(probs)
        return out.float(), probs

    @staticmethod
    def backward(ctx, d_out, _d_probs):
        qa, ka, va, probs = ctx.saved_tensors
        G, Tq, Tk, h, dk, p, seed, stream, adt, has_bias = ctx.meta
        d = h * dk
        g = d_out.float().contiguous()
        dq = torch.zeros(G * Tq, d, device=g.device)
        dk_ = torch.zeros(G * Tk, d, device=g.device)
        dv = torch.zeros(G * Tk, d, device=g.device)
        dbias = torch.zeros(G, h, Tq, Tk, device=g.device) if has_bias else None
        K.attention_bwd(qa, ka, va, probs, g, dq, dk_, dv, dtype=adt, G=G, Tq=Tq, Tk=Tk, h=h, dk=dk, ldq=d, ldk=d, ldv=d, ldd=d,
                        ldgq=d, ldgk=d, ldgv=d, dbias=dbias, p=p, seed=seed, stream_id=stream)
        return dq, dk_, dv, dbias, None, None, None, None, None, None, None, None, None, None


def attention(q, k, v, *, G, Tq, Tk, h, bias=None, key_valid=None, causal_T=0, p=0.0, training=False, precision="bf16"):
    """Returns (out fp32 [G*Tq, d], probs fp32 [G, h, Tq, Tk] after dropout scaling is NOT applied - raw softmax)."""
    _need_cuda(q, "attention")
    p = float(p) if training else 0.0
    seed, stream = sampler.next_dropout_stream() if p > 0 else (0, 0)
    return _AttentionFn.apply(q, k, v, bias, key_valid, G, Tq, Tk, h, causal_T, p, seed, stream, precision)


# ---- box geometry (relation_transformer.py:179-183,196-256,283-286) ---------------------------------------------------
class _BoxBiasFn(torch.autograd.Function):
    """log(max(relu(WG_h . emb(i, j) + b_h), 1e-6)) for all heads straight from the boxes (the [B,N,N,64] embedding is never
    materialised); gradients to the (masked) WG weights and biases."""

    @staticmethod
    def forward(ctx, boxes, wg_w, wg_b, trig):
        B, N = boxes.shape[:2]
        h = wg_w.shape[0]
        bx = boxes.detach().float().contiguous()
        bias = torch.empty(B, h, N, N, device=boxes.device)
        K.box_bias_fwd(bx, wg_w.detach().float().contiguous(), wg_b.detach().float().contiguous(), bias, B=B, N=N, h=h, trig=trig)
        ctx.save_for_backward(bx, bias)
        ctx.meta = (B, N, h, trig, wg_w.shape)
        return bias

    @staticmethod
    def backward(ctx, dbias):
        bx, bias = ctx.saved_tensors
        B, N, h, trig, wshape = ctx.meta
        dw = torch.zeros(wshape, device=bx.device)
        db = torch.zeros(h, device=bx.device)
        K.box_bias_bwd(bx, bias, dbias.float().contiguous(), dw, db, B=B, N=N, h=h, trig=trig)
        return None, dw, db, None


def box_bias(boxes, wg_w, wg_b, trig=True):
    _need_cuda(boxes, "box geometry bias")
    return _BoxBiasFn.apply(boxes, wg_w, wg_b, trig)


class _LogClampFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, lo):
        xc = x.detach().float().contiguous()
        ctx.save_for_backward(xc)
        ctx.lo = lo
        return K.log_clamp(xc, lo=lo)

    @staticmethod
    def backward(ctx, dy):
        (xc,) = ctx.saved_tensors
        return K.log_clamp(xc, lo=ctx.lo, dy=dy.float().contiguous()), None


def log_clamp(x, lo=1e-6):
    _need_cuda(x, "log_clamp")
    return _LogClampFn.apply(x, lo)


# ---- generator: log_softmax (transformer.py:405-413) -------------------------------------------------------------------
class _LogSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits):
        V = logits.shape[-1]
        x2 = logits.reshape(-1, V).float().contiguous()
        lp = torch.empty_like(x2)
        K.logsoftmax_nll(x2, logprobs=lp)
        ctx.save_for_backward(lp)
        ctx.shape = logits.shape
        return lp.reshape(logits.shape)

    @staticmethod
    def backward(ctx, dy):
        (lp,) = ctx.saved_tensors
        return K.logsoftmax_bwd(lp, dy.reshape(lp.shape).float().contiguous()).reshape(ctx.shape)


def log_softmax(logits):
    _need_cuda(logits, "log_softmax")
    return _LogSoftmaxFn.apply(logits)


SQRT = math.sqrt
