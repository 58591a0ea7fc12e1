"""Mask samplers of the supermask layers — drop-in for sparse_caption/pruning/sampler.py.

``bernoulli_sample_sigmoid`` / ``rounding_sigmoid`` keep the reference's signatures (sampler.py:43-66).  On CUDA
tensors the forward runs in sc_apply_mask (Philox Bernoulli / bit-exact binarization); the backward is the
straight-through estimator of the reference: grad wrt logits = grad_out * sigmoid'(S), or grad_out itself when
``bypass_sigmoid_grad`` (sampler.py:15-17, 32-34).
"""
import torch

from . import kernels as K

_state = {"seed": 0, "stream": 0, "drop": 0, "uniforms": None}


def set_mask_seed(seed: int) -> None:
    """Seed of the Philox stream behind every Bernoulli mask.  All data-parallel ranks must use the same seed so
    that they draw the same mask (the reference draws ONE mask per layer per step, masked_layer.py:97)."""
    _state["seed"] = int(seed)
    _state["stream"] = 0


def next_dropout_stream():
    """(seed, stream_id) of the next dropout site (activation / attention dropout of the module-level path); a stream
    space disjoint from the mask streams."""
    _state["drop"] += 1
    return _state["seed"] + 1, (1 << 40) + _state["drop"]


def inject_uniforms(mapping) -> None:
    """Parity tests: ``{mask_logits Parameter: uniforms tensor}`` - train-mode supermasks then use ``u < sigmoid(S)`` with the
    given uniforms instead of the Philox stream (the reference side patches torch.bernoulli the same way).  None clears."""
    _state["uniforms"] = None if mapping is None else {id(k): v for k, v in mapping.items()}


def injected_uniforms(logits):
    u = _state["uniforms"]
    if u is None or logits is None:
        return None
    return u.get(id(logits))


def next_mask_stream():
    """(seed, stream_id) for the next sampled mask; stream ids advance once per sampled tensor."""
    _state["stream"] += 1
    return _state["seed"], _state["stream"]


class _SampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, mode, bypass):
        if not logits.is_cuda:
            raise RuntimeError("mask sampling runs on the GPU only (no CPU fallback); move the module to CUDA")
        seed, stream = next_mask_stream() if mode == K.MASK_BERNOULLI else (0, 0)
        ones = torch.ones_like(logits)
        ctx.save_for_backward(logits)
        ctx.bypass = bypass
        return K.apply_mask(ones, logits.detach().contiguous(), mode, seed=seed, stream_id=stream)

    @staticmethod
    def backward(ctx, grad):
        (logits,) = ctx.saved_tensors
        if ctx.bypass:
            return grad, None, None
        p = torch.sigmoid(logits)
        return grad * p * (1 - p), None, None


def bernoulli_sample_sigmoid(logits, bypass_sigmoid_grad=False):
    """Stochastic Bernoulli(sigmoid(logits)) sample, straight-through gradient (sampler.py:43-54)."""
    return _SampleFn.apply(logits, K.MASK_BERNOULLI, bypass_sigmoid_grad)


def rounding_sigmoid(logits, bypass_sigmoid_grad=False):
    """Deterministic rint(sigmoid(logits)), straight-through gradient (sampler.py:57-66)."""
    return _SampleFn.apply(logits, K.MASK_ROUND, bypass_sigmoid_grad)
