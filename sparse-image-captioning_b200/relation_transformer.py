"""Drop-in model classes for the ORT / ACORT hot path — mirrors the public surface of
sparse_caption/models/{__init__,caption_model,transformer,relation_transformer,relation_transformer_prune}.py.

* ``get_model("relation_transformer")(config)`` / ``get_model("relation_transformer_prune")(config)`` build modules whose
  parameter names equal the reference's (``model.encoder.layers.0.self_attn.WGs.3.weight_pruning_mask`` ...), so
  reference checkpoints load with ``strict=True``.
* ``model(att_feats=, boxes=, seqs=, att_masks=)`` returns teacher-forcing log-probs ``[B*S, T, V]``;
  ``model(att_feats=, boxes=, att_masks=, opt=, mode="sample")`` returns ``(seq [B,b,L] int64, seq_logprobs [B,b,L])``;
  ``get_logprobs_state`` / ``batch_beam_search`` are kept as entry points.
* Every arithmetic step runs in the CUDA kernels behind include/sc_b200.h (OrtEngine for decoding, OrtTrainer for
  teacher forcing / training).  The module tree only OWNS the parameters; there is no PyTorch forward to fall back to,
  and calling a model that lives on the CPU raises.
"""
import logging
import math
from argparse import ArgumentParser, _ArgumentGroup
from typing import Any, Union

import torch
from torch import nn

from . import prune
from .engine import ModelCfg, OrtEngine
from .masked_layer import MaskedEmbedding, MaskedLinear
from .prune import PruningMixin

logger = logging.getLogger(__name__)

MODEL_REGISTRY = {}


def register_model(name):
    """Same decorator contract as sparse_caption/models/__init__.py:16-36."""

    def deco(cls):
        if name in MODEL_REGISTRY:
            raise ValueError(f"Cannot register duplicate model: `{name}`.")
        MODEL_REGISTRY[name.lower()] = cls
        return cls

    return deco


def get_model(name: str) -> Any:
    try:
        return MODEL_REGISTRY[name.lower()]
    except KeyError:
        raise ValueError(f"Model specified `{name}` is invalid. Available options are: \n" + "\n".join(MODEL_REGISTRY))


# ---- parameter containers (names follow the reference module tree) ------------------------------------------------
class LayerNorm(nn.Module):
    def __init__(self, features, eps=1e-6):
        super().__init__()
        self.a_2 = nn.Parameter(torch.ones(features))
        self.b_2 = nn.Parameter(torch.zeros(features))
        self.eps = eps


class SublayerConnection(nn.Module):
    def __init__(self, size, dropout):
        super().__init__()
        self.norm = LayerNorm(size)
        self.dropout = nn.Dropout(dropout)


def _linear(mask, i, o):
    return MaskedLinear(i, o, mask[0], mask[1]) if mask else nn.Linear(i, o)


class _Attention(nn.Module):
    def __init__(self, mask, h, d_model, share_att, box, trig=True, dropout=0.1):
        super().__init__()
        assert d_model % h == 0
        assert share_att in (None, "kv", "qk"), f"Invalid `share_att`: {share_att}"
        self.h, self.d_k, self.share_att = h, d_model // h, share_att
        self.linears = nn.ModuleList([_linear(mask, d_model, d_model) for _ in range(3 if share_att else 4)])
        if box:
            self.trigonometric_embedding = trig
            self.dim_g = 64 if trig else 4
            self.WGs = nn.ModuleList([_linear(mask, self.dim_g, 1) for _ in range(h)])
        else:
            self.self_attention = False
        self.dropout = nn.Dropout(p=dropout)


class PositionwiseFeedForward(nn.Module):
    def __init__(self, mask, d_model, d_ff, dropout):
        super().__init__()
        self.w_1 = _linear(mask, d_model, d_ff)
        self.w_2 = _linear(mask, d_ff, d_model)
        self.dropout = nn.Dropout(dropout)


class EncoderLayer(nn.Module):
    def __init__(self, mask, cfg, dropout):
        super().__init__()
        d = cfg.d_model
        self.self_attn = _Attention(mask, cfg.num_heads, d, cfg.share_att_encoder, True, not cfg.no_box_trigonometric_embedding, dropout)
        self.feed_forward = PositionwiseFeedForward(mask, d, cfg.dim_feedforward, dropout)
        self.sublayer = nn.ModuleList([SublayerConnection(d, dropout) for _ in range(2)])
        self.size = d


class DecoderLayer(nn.Module):
    def __init__(self, mask, cfg, dropout):
        super().__init__()
        d = cfg.d_model
        self.size = d
        self.self_attn = _Attention(mask, cfg.num_heads, d, cfg.share_att_decoder, False, dropout=dropout)
        self.self_attn.self_attention = True
        self.src_attn = _Attention(mask, cfg.num_heads, d, cfg.share_att_decoder, False, dropout=dropout)
        self.feed_forward = PositionwiseFeedForward(mask, d, cfg.dim_feedforward, dropout)
        self.sublayer = nn.ModuleList([SublayerConnection(d, dropout) for _ in range(3)])


class _Stack(nn.Module):
    """Encoder / Decoder: N layers (shared according to ``share_layer``) + final norm (relation_transformer.py:77-96)."""

    def __init__(self, make_layer, N, size, share_layer=None):
        super().__init__()
        if share_layer:
            if not isinstance(share_layer, (tuple, list)):
                raise TypeError(f"`share_layer` must be a tuple or list, saw {type(share_layer)}")
            uniq = [make_layer() for _ in range(len(set(share_layer)))]
            layers = [uniq[i] for i in share_layer]
        else:
            layers = [make_layer() for _ in range(N)]
        self.layers = nn.ModuleList(layers)
        self.norm = LayerNorm(size)


class InputEmbedding(nn.Module):
    def __init__(self, mask, d_model, vocab):
        super().__init__()
        self.lut = MaskedEmbedding(vocab, d_model, mask[0], mask[1]) if mask else nn.Embedding(vocab, d_model)
        self.d_model = d_model


class PositionalEncoding(nn.Module):
    def __init__(self, d_model, dropout, max_len=5000):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len).unsqueeze(1).float()
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0))


class OutputEmbedding(nn.Module):
    def __init__(self, mask, d_model, vocab):
        super().__init__()
        self.proj = _linear(mask, d_model, vocab)


class EncoderDecoder(nn.Module):
    def __init__(self, encoder, decoder, tgt_embed, generator):
        super().__init__()
        self.encoder, self.decoder, self.tgt_embed, self.generator = encoder, decoder, tgt_embed, generator


class PrunedEncoderDecoder(PruningMixin, EncoderDecoder):
    pass


# ---- models -------------------------------------------------------------------------------------------------------
class _OrtBase(nn.Module):
    """CaptionModel.forward dispatch (caption_model.py:24-28) + the kernel-backed entry points."""

    MASKED = False

    def _init_common(self, config):
        self.config = config
        self.cfg = ModelCfg(config)
        for f in ModelCfg.FIELDS:
            setattr(self, {"eos_token_id": "eos_idx", "bos_token_id": "bos_idx", "unk_token_id": "unk_idx",
                           "pad_token_id": "pad_idx", "max_seq_length": "seq_length"}.get(f, f), getattr(self.cfg, f))
        self.drop_prob_src = getattr(config, "drop_prob_src", 0.5) if not isinstance(config, dict) else config.get("drop_prob_src", 0.5)
        self.precision = "bf16"
        self._engine = None
        self._engine_key = None
        self._trainer = None

    def make_model(self, mask, dropout):
        c = self.cfg
        enc = _Stack(lambda: EncoderLayer(mask, c, dropout), c.num_layers, c.d_model, c.share_layer_encoder)
        dec = _Stack(lambda: DecoderLayer(mask, c, dropout), c.num_layers, c.d_model, c.share_layer_decoder)
        tgt = nn.Sequential(InputEmbedding(mask, c.d_model, c.vocab_size), PositionalEncoding(c.d_model, dropout))
        gen = OutputEmbedding(mask, c.d_model, c.vocab_size)
        if mask:
            model = PrunedEncoderDecoder(mask_type=mask[0], mask_freeze_scope="", encoder=enc, decoder=dec, tgt_embed=tgt, generator=gen)
            weights = model.all_weights(named=False)
        else:
            model = EncoderDecoder(enc, dec, tgt, gen)
            weights = list(model.parameters())
        self.att_embed = nn.Sequential(_linear(mask, c.att_feat_size, c.d_model), nn.ReLU(), nn.Dropout(self.drop_prob_src))
        for p in weights:  # Glorot / fan_avg on the inner model only (relation_transformer.py:333-337)
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        self.model = model
        self.dropout_p = dropout

    # -- dispatch ----------------------------------------------------------------------------------------------
    def forward(self, *args, **kwargs):
        mode = kwargs.pop("mode", "forward")
        return getattr(self, "_" + mode)(*args, **kwargs)

    def _device(self):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("the B200 captioning path runs on CUDA only (no CPU fallback): call model.cuda() first")
        return dev

    def _effective_state_dict(self):
        sd = self.state_dict()
        if self.MASKED:
            return prune.fold_masks(sd, self.mask_type)
        return sd

    def _get_engine(self):
        dev = self._device()
        key = (tuple(p._version for p in self.parameters()), self.precision, self.training if self.MASKED else False)
        if self._engine is None or self._engine_key != key:
            self._engine = OrtEngine(self._effective_state_dict(), self.cfg, precision=self.precision, device=dev,
                                     no_history=self.MASKED and getattr(self, "compat_no_history", False))
            self._engine_key = key
        return self._engine

    # -- inference ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _sample(self, att_feats, boxes, att_masks=None, opt=None, **kwargs):
        """relation_transformer.py:390-396 -> OrtEngine.sample (encoder + KV-cached beam / greedy search).
        Masks are binarized (eval semantics) when the pruned class samples; the reference's train-mode Bernoulli
        sampling during SCST rollouts is not reproduced."""
        opt = dict(opt or {})
        return self._get_engine().sample(att_feats, boxes, att_masks, opt)

    @torch.no_grad()
    def _forward(self, att_feats, boxes, seqs, att_masks=None, **kwargs):
        """Teacher-forcing log-probs [B*S, T, V] (relation_transformer.py:368-372), forward only.  Training steps go
        through ``trainer()`` (explicit backward kernels); autograd through this call is not supported."""
        from . import kernels as K
        tr = self.trainer()
        tr.training = self.training
        B, N = att_feats.shape[:2]
        R = seqs.shape[0]
        T = seqs.shape[1] - 1
        ws = tr._get_ws(B, N, R // B, T, att_masks is not None)
        tr.step_id += 1
        masks = torch.ones(seqs.shape, device=seqs.device)
        tr.load_batch(ws, att_feats, boxes, seqs, masks, att_masks)
        logits = tr.forward(ws)
        lp = torch.empty(ws.MD, tr.Vp, device=logits.device)
        K.logsoftmax_nll(logits, logprobs=lp)
        return lp[:, : self.cfg.vocab_size].reshape(R, T, -1)

    def trainer(self, **kw):
        """The kernel-backed training engine bound to this module's parameters (created on first use).
        ``trainer().train_step(...)`` replaces loss.backward() + clip_gradient + optimizer.step() of the reference
        loop (scripts/train_n_prune_transformer.py:136-153); ``sync_from_trainer()`` copies the result back."""
        from .trainer import OrtTrainer
        if self._trainer is None:
            self._trainer = OrtTrainer(self.state_dict(), self.cfg, mask_type=self.mask_type if self.MASKED else None,
                                       precision=self.precision, device=self._device(), dropout=self.dropout_p,
                                       drop_prob_src=self.drop_prob_src, **kw)
        return self._trainer

    @torch.no_grad()
    def sync_from_trainer(self):
        self.load_state_dict(self._trainer.state_dict(), strict=False)

    def get_logprobs_state(self, it, memory, mask, state):
        raise NotImplementedError("step-wise decoding is fused on the device: use mode='sample' (OrtEngine.decode); "
                                  "the per-step entry point of the reference has no host-visible state here")

    def batch_beam_search(self, init_state, init_logprobs, *args, **kwargs):
        raise NotImplementedError("beam search runs inside OrtEngine.decode (sc_beam_step); use mode='sample'")

    @staticmethod
    def clip_att(att_feats, att_masks):
        if att_masks is not None:
            max_len = att_masks.data.long().sum(1).max()
            att_feats = att_feats[:, :max_len].contiguous()
            att_masks = att_masks[:, :max_len].contiguous()
        return att_feats, att_masks

    @staticmethod
    def subsequent_mask(size):
        return torch.triu(torch.ones((1, size, size)), diagonal=1).eq(0)

    @staticmethod
    def add_argparse_args(parser: Union[_ArgumentGroup, ArgumentParser]):
        parser.add_argument("--d_model", type=int, default=512)
        parser.add_argument("--dim_feedforward", type=int, default=2048)
        parser.add_argument("--num_layers", type=int, default=6)
        parser.add_argument("--num_heads", type=int, default=8)
        parser.add_argument("--drop_prob_src", type=float, default=0.5)
        parser.add_argument("--att_feat_size", type=int, default=2048)
        parser.add_argument("--share_att_encoder", type=str, default=None)
        parser.add_argument("--share_att_decoder", type=str, default=None)
        parser.add_argument("--share_layer_encoder", type=str, default=None)
        parser.add_argument("--share_layer_decoder", type=str, default=None)
        parser.add_argument("--no_box_trigonometric_embedding", action="store_true")


@register_model("relation_transformer")
class RelationTransformerModel(_OrtBase):
    def __init__(self, config):
        super().__init__()
        self._init_common(config)
        self.make_model(None, 0.1)


@register_model("relation_transformer_prune")
class RelationTransformerPruneModel(PruningMixin, _OrtBase):
    MASKED = True

    def __init__(self, config):
        get = (lambda k, d=None: config.get(k, d)) if isinstance(config, dict) else (lambda k, d=None: getattr(config, k, d))
        super().__init__(mask_type=get("prune_type"), mask_freeze_scope=get("prune_mask_freeze_scope", ""))
        self._init_common(config)
        self.make_model((get("prune_type"), get("prune_supermask_init", 5.0)), 0.1 / 3)

    @staticmethod
    def add_argparse_args(parser):
        _OrtBase.add_argparse_args(parser)
        PruningMixin.add_argparse_args(parser)
