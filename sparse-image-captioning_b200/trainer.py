"""Supermask (SMP) training step of the ORT captioner on the B200 kernels.

Replaces the reference's hot loop body (scripts/train_n_prune_transformer.py:132-156):
    model(**data) -> LanguageModelCriterion -> + compute_sparsity_loss -> backward -> clip_gradient(0.1) -> Adam
for `relation_transformer_prune` (sparse_caption/models/relation_transformer_prune.py) and, with mask_type=None, for the
dense `relation_transformer`.  The encoder runs once per image; the S captions of an image share its memory K/V
(the reference repeats the memory S times before the cross K/V projections, relation_transformer.py:63-66).

All parameters live in two flat fp32 buffers (weights+biases+norms | mask logits) with views per reference parameter
name, so the optimizer is two fused kernels and the data-parallel gradient exchange is an NCCL all-reduce over flat
gradient buffers (mask-logit gradients included).  Forward and backward are explicit sequences of kernel launches;
no autograd graph is built.
"""
import math
from typing import Dict, Optional

import torch

from . import kernels as K
from .engine import ModelCfg


def _round_up(n, m):
    return (n + m - 1) // m * m


class OrtTrainer:
    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: ModelCfg, *, mask_type: Optional[str] = "supermask",
                 precision="bf16", device="cuda", dropout=0.1 / 3, drop_prob_src=0.5, bypass_sigmoid_grad=False, seed=0,
                 mask_init_value=5.0, uniforms: Optional[Dict[str, torch.Tensor]] = None, use_graph=False, fused_st=True):
        assert cfg.share_att_encoder is None and cfg.share_att_decoder is None and not cfg.share_layer_encoder \
            and not cfg.share_layer_decoder, "the fused engine lays out the ORT stack; ACORT weight sharing trains through ModuleTrainer (model.trainer() picks it)"
        self.cfg = cfg
        self.dev = K.lib.resolve_device(device)
        self.adt = torch.bfloat16 if precision == "bf16" else torch.float32
        self.mask_type = mask_type
        self.p_drop, self.p_src = float(dropout), float(drop_prob_src)
        self.bypass = bool(bypass_sigmoid_grad)
        self.seed = int(seed)
        self.step_id = 0
        self.training = True
        # CUDA-graph mode: the whole step (forward, backward, optimizer) is captured once per batch shape and replayed;
        # everything that changes per step (Philox seeds, lr, Adam bias corrections, sparsity-loss scale) is read from
        # small device buffers that the host rewrites before each replay (sc_b200.h: seed pointers, `dyn`)
        self.use_graph = bool(use_graph)
        # straight-through epilogue inside the optimizer kernel (sc_adam_clip_st): the backward leaves dWm = dL/d(W.m) in flat_gw,
        # which is ALL a data-parallel run exchanges (half the bytes of dW + dS); dW = dWm.m and dS = dWm.W.sigmoid'(S) are formed
        # per element inside the update with the step's mask sample regenerated.  False: the weight-gradient epilogue writes dW and
        # dS (needed by the ZeRO-style ShardedExchange, whose shards cut through tensors).
        self.fused_st = bool(fused_st)
        self._materialized = False
        self._st_desc = {}
        self._seeds_host = torch.zeros(3, dtype=torch.int64).pin_memory() if self.use_graph else None
        self._seeds_dev = torch.zeros(3, dtype=torch.int64, device=self.dev)
        self._dyn_host = torch.zeros(8, dtype=torch.float32).pin_memory() if self.use_graph else None
        self._dyn_dev = torch.zeros(8, dtype=torch.float32, device=self.dev)
        d, L = cfg.d_model, cfg.num_layers
        sd = state_dict
        # ---- parameter layout: fused groups are contiguous so that [q;k;v] / [k;v] / WG blocks are single views ----
        # the one-element WG biases of ALL encoder layers come first, contiguous: the geometry bias of every layer is then
        # one kernel (sc_box_bias_all) over a single [L*h] bias vector
        order = [f"model.encoder.layers.{i}.self_attn.WGs.{j}.bias" for i in range(L) for j in range(cfg.num_heads)]
        order += ["att_embed.0.weight", "att_embed.0.bias"]
        for i in range(L):
            p = f"model.encoder.layers.{i}"
            order += [f"{p}.self_attn.linears.{j}.weight" for j in range(3)] + [f"{p}.self_attn.linears.{j}.bias" for j in range(3)]
            order += [f"{p}.self_attn.linears.3.weight", f"{p}.self_attn.linears.3.bias"]
            order += [f"{p}.self_attn.WGs.{j}.weight" for j in range(cfg.num_heads)]
            order += [f"{p}.feed_forward.w_1.weight", f"{p}.feed_forward.w_1.bias", f"{p}.feed_forward.w_2.weight", f"{p}.feed_forward.w_2.bias"]
            order += [f"{p}.sublayer.{j}.norm.{ab}" for j in range(2) for ab in ("a_2", "b_2")]
        order += ["model.encoder.norm.a_2", "model.encoder.norm.b_2"]
        for i in range(L):
            p = f"model.decoder.layers.{i}"
            for att in ("self_attn", "src_attn"):
                order += [f"{p}.{att}.linears.{j}.weight" for j in range(3)] + [f"{p}.{att}.linears.{j}.bias" for j in range(3)]
                order += [f"{p}.{att}.linears.3.weight", f"{p}.{att}.linears.3.bias"]
            order += [f"{p}.feed_forward.w_1.weight", f"{p}.feed_forward.w_1.bias", f"{p}.feed_forward.w_2.weight", f"{p}.feed_forward.w_2.bias"]
            order += [f"{p}.sublayer.{j}.norm.{ab}" for j in range(3) for ab in ("a_2", "b_2")]
        order += ["model.decoder.norm.a_2", "model.decoder.norm.b_2", "model.tgt_embed.0.lut.weight",
                  "model.generator.proj.weight", "model.generator.proj.bias"]
        self.names = order
        # the generator is stored with its vocabulary padded to a multiple of 8 rows (zero weights, bias -1e9, masked
        # out) so that dlogits can be a 16-byte-aligned bf16 GEMM operand for any V (e.g. the 771-token radix vocab)
        V = cfg.vocab_size
        self.Vp = K.pad8(V)
        self._pad_rows = {"model.generator.proj.weight": self.Vp, "model.generator.proj.bias": self.Vp}

        def alloc(k):
            if k in self._pad_rows:
                return sd[k].numel() // V * self.Vp
            return sd[k].numel()

        def layout(keys):
            # every view 16-byte aligned, except that the h one-element WG biases of a layer are packed into one [h] block
            offs, n = {}, 0
            for k in keys:
                offs[k] = n
                n += alloc(k)
                if not (".WGs." in k and k.endswith(".bias") and not k.endswith(f".WGs.{cfg.num_heads - 1}.bias")):
                    n = _round_up(n, 4)
            return offs, n

        offs, n = layout(order)
        self.flat_w = torch.zeros(n, device=self.dev)
        self.flat_gw = torch.zeros(n, device=self.dev)
        self.p = {k: self.flat_w[offs[k]: offs[k] + sd[k].numel()].view(sd[k].shape) for k in order}
        self.g = {k: self.flat_gw[offs[k]: offs[k] + sd[k].numel()].view(sd[k].shape) for k in order}
        for k in order:
            self.p[k].copy_(sd[k])
        self._offs_w = offs
        if self.Vp != V:
            ob = offs["model.generator.proj.bias"]
            self.flat_w[ob + V: ob + self.Vp] = -1e9
        # mask logits for every 2-D weight
        self.masked = [k for k in order if k.endswith(".weight") and sd[k].dim() == 2] if mask_type else []
        offs_s, n = layout(self.masked)
        self.flat_s = torch.zeros(max(n, 4), device=self.dev)
        self.flat_gs = torch.zeros(max(n, 4), device=self.dev)
        self.n_logits = sum(sd[k].numel() for k in self.masked)
        self.s = {k: self.flat_s[offs_s[k]: offs_s[k] + sd[k].numel()].view(sd[k].shape) for k in self.masked}
        self.gs = {k: self.flat_gs[offs_s[k]: offs_s[k] + sd[k].numel()].view(sd[k].shape) for k in self.masked}
        for k in self.masked:
            mk = k + "_pruning_mask"
            if mk in sd:
                self.s[k].copy_(sd[mk])
            else:
                self.s[k].fill_(mask_init_value)
        self._offs_s = offs_s
        # padding elements of flat_s must never count as "kept" in the sparsity statistics
        pad = torch.ones_like(self.flat_s, dtype=torch.bool)
        for k in self.masked:
            pad[offs_s[k]: offs_s[k] + sd[k].numel()] = False
        self.flat_s[pad] = -1.0
        self._s_pad_mask = pad
        self._s_pad_idx = pad.nonzero().view(-1)  # index form: index_fill_ is CUDA-graph capturable
        self.stream_of = {k: i + 1 for i, k in enumerate(order)}
        # injected uniforms (parity tests): same layout as flat_s
        self.flat_u = None
        if uniforms is not None:
            self.flat_u = torch.zeros_like(self.flat_s)
            for k in self.masked:
                self.flat_u[offs_s[k]: offs_s[k] + sd[k].numel()].view(sd[k].shape).copy_(uniforms[k])
        pe = sd.get("model.tgt_embed.1.pe")
        from .engine import _positional_encoding
        self.pe = (pe[0, : cfg.max_seq_length + 2].to(self.dev) if pe is not None
                   else _positional_encoding(d, cfg.max_seq_length + 2, self.dev)).float().contiguous()
        # optimizer state
        self.m_w, self.v_w = torch.zeros_like(self.flat_w), torch.zeros_like(self.flat_w)
        self.m_s, self.v_s = torch.zeros_like(self.flat_s), torch.zeros_like(self.flat_s)
        self.opt_step = 0
        self.sp_out = torch.zeros(3, device=self.dev)
        self.sp_count = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self._ws = {}
        self.premask = True   # False: mask inside the GEMM operand prologue (K1 fused variant)
        self._wm, self._wmT, self._wm_step = {}, {}, {}
        self._premask_desc = None
        # timing diagnostic (scripts/profile_train.py SC_SKIP_WGRAD=1: wrong gradients): the main chain alone - measured 4.95 of the
        # 5.35 ms step, i.e. the weight-gradient side stream costs 0.4 ms and the step is bound by the ~290 dependent launches
        self._diag_skip_wgrad = False
        self.pdl_mask = 3            # sc_set_pdl mask while the step is launched / captured (see train_step)
        self.wgrad_ring = 4          # 1: weight gradients stay on the main stream
        import os
        self.fuse_hmask = os.environ.get("SC_NO_HMASK") != "1"      # ff2's dX GEMM prepares ff1's gradient operand
        # attention backward emitting bf16 operands + bias gradients itself (sc_attention_bwd_bf16out): parity-tested, but
        # measured SLOWER (5.66 vs 5.58 ms/step: ~640k same-address atomics for the bias sums) - off unless SC_ATTN16=1
        self.fuse_attn_bwd = os.environ.get("SC_ATTN16") == "1"
        self._side = None
        self._ost = None
        # decoder-side Adam on its own stream underneath the encoder backward: measured slower (5.42 vs 5.35 ms/step, the
        # HBM-bound update slows the backward kernels more than it hides) - off unless SC_OPT_OVERLAP=1
        self.overlap_opt = os.environ.get("SC_OPT_OVERLAP") == "1"

    # ---------------------------------------------------------------------------------------------------------
    def _group(self, table, first, count):
        """Contiguous view over `count` consecutive same-shape parameters starting at `first` (e.g. [q;k;v])."""
        shp = table[first].shape
        offs = self._offs_w if table is self.p or table is self.g else self._offs_s
        flat = {id(self.p): self.flat_w, id(self.g): self.flat_gw, id(self.s): self.flat_s, id(self.gs): self.flat_gs}[id(table)]
        n = table[first].numel()
        if first in self._pad_rows and count == 1:
            rows = self._pad_rows[first]
            return flat[offs[first]: offs[first] + n // shp[0] * rows].view((rows,) + tuple(shp[1:]))
        return flat[offs[first]: offs[first] + n * count].view((shp[0] * count,) + tuple(shp[1:]))

    def _u(self, first, count=1):
        if self.flat_u is None:
            return None
        n = self.s[first].numel()
        shp = self.s[first].shape
        if first in self._pad_rows and count == 1:
            rows = self._pad_rows[first]
            return self.flat_u[self._offs_s[first]: self._offs_s[first] + n // shp[0] * rows].view((rows,) + tuple(shp[1:]))
        return self.flat_u[self._offs_s[first]: self._offs_s[first] + n * count].view((shp[0] * count,) + tuple(shp[1:]))

    def mask_mode(self):
        if not self.mask_type:
            return K.MASK_NONE
        if self.mask_type == "supermask":
            if not self.training:
                return K.MASK_ROUND
            return K.MASK_UNIFORM if self.flat_u is not None else K.MASK_BERNOULLI
        return K.MASK_RAW

    def _sd(self, i):
        """Philox seed i (0 masks, 1 dropout, 2 attention dropout): an immediate value, or in graph mode a device pointer
        (bit 63 set) to the seed word the host rewrites every step."""
        if self.use_graph:
            return (1 << 63) | (self._seeds_dev.data_ptr() + 8 * i)
        return self.seed + i

    def _step_base(self):
        return 0 if self.use_graph else self.step_id * 4096

    def _mask_args(self, wname, count=1):
        """(weight view, logits view, mode, uniforms view, seed, stream) of a (fused) masked weight."""
        W = self._group(self.p, wname, count)
        if not self.mask_type:
            return W, None, K.MASK_NONE, None, 0, 0
        return (W, self._group(self.s, wname, count), self.mask_mode(), self._u(wname, count), self._sd(0),
                self._step_base() + self.stream_of[wname])

    def _premask_keys(self):
        """(first weight name, fused count) of every linear the teacher-forcing forward runs, in launch order."""
        L = self.cfg.num_layers
        keys = [("att_embed.0.weight", 1)]
        for l in range(L):
            p = f"model.encoder.layers.{l}"
            keys += [(f"{p}.self_attn.linears.0.weight", 3), (f"{p}.self_attn.linears.3.weight", 1),
                     (f"{p}.feed_forward.w_1.weight", 1), (f"{p}.feed_forward.w_2.weight", 1)]
        keys += [(f"model.decoder.layers.{l}.src_attn.linears.1.weight", 2) for l in range(L)]
        for l in range(L):
            p = f"model.decoder.layers.{l}"
            keys += [(f"{p}.self_attn.linears.0.weight", 3), (f"{p}.self_attn.linears.3.weight", 1),
                     (f"{p}.src_attn.linears.0.weight", 1), (f"{p}.src_attn.linears.3.weight", 1),
                     (f"{p}.feed_forward.w_1.weight", 1), (f"{p}.feed_forward.w_2.weight", 1)]
        keys.append(("model.generator.proj.weight", 1))
        return keys

    def _premask_all(self):
        """W (.) m (bf16) and its transpose for EVERY linear of the step in one launch (sc_apply_mask_batched): the forward
        operand, the dX operand and the weight-gradient epilogue all see the same Philox sample."""
        if self._premask_desc is None:
            items = []
            for key in self._premask_keys():
                wname, count = key
                W, S, mode, U, seed, stream = self._mask_args(wname, count)
                wm = self._wm[key] = torch.empty(W.shape, device=self.dev, dtype=torch.bfloat16)
                wT = self._wmT[key] = torch.empty(W.shape[1], W.shape[0], device=self.dev, dtype=torch.bfloat16)
                items.append((W, S, U, wm, wT, self.stream_of[wname]))
            self._premask_desc = K.mask_descriptors(items, self.dev) + ([k for k in self._premask_keys()],)
        desc, tiles, keys = self._premask_desc
        K.apply_mask_batched(desc, tiles, self.mask_mode(), seed=self._sd(0), stream_base=self._step_base())
        for key in keys:
            self._wm_step[key] = (self.step_id, self.training)

    def _drop_stream(self, site):
        return self._step_base() + 2048 + site

    # ---------------------------------------------------------------------------------------------------------
    def _get_ws(self, B, N, S, T, masked):
        key = (B, N, S, T, masked)
        if key in self._ws:
            return self._ws[key]
        c, dev, adt = self.cfg, self.dev, self.adt
        d, ff, V, L, h, F = c.d_model, c.dim_feedforward, c.vocab_size, c.num_layers, c.num_heads, c.att_feat_size
        ME, R = B * N, B * S
        MD = R * T
        Mp = K.pad8(max(ME, MD))
        f32 = dict(device=dev, dtype=torch.float32)
        a = dict(device=dev, dtype=adt)
        ws = type("TrainWs", (), {})()
        ws.B, ws.N, ws.S, ws.T, ws.ME, ws.MD, ws.R, ws.Mp = B, N, S, T, ME, MD, R, Mp
        ws.att_in = torch.zeros(ME, F, **f32)
        ws.att_a = torch.zeros(ME, F, **a) if adt != torch.float32 else ws.att_in
        ws.boxes = torch.zeros(B, N, 4, **f32)
        ws.att_mask = torch.ones(B, N, **f32) if masked else None
        ws.tokens = torch.zeros(MD, device=dev, dtype=torch.int32)
        ws.targets = torch.zeros(MD, device=dev, dtype=torch.int32)
        ws.tok_w = torch.zeros(MD, **f32)
        ws.key_valid = torch.zeros(R, T, **f32)
        ws.inv_norm = torch.zeros(1, **f32)
        ws.loss_sum = torch.zeros(1, **f32)
        ws.xe = [torch.zeros(ME, d, **f32) for _ in range(2 * L + 1)]
        ws.e_xn1 = [torch.zeros(ME, d, **a) for _ in range(L)]
        ws.e_qkv = [torch.zeros(ME, 3 * d, **a) for _ in range(L)]
        ws.e_bias_all = torch.zeros(L, B, h, N, N, **f32)
        ws.e_bias = [ws.e_bias_all[l] for l in range(L)]
        ws.e_probs = [torch.zeros(B, h, N, N, **f32) for _ in range(L)]
        ws.e_att = [torch.zeros(ME, d, **a) for _ in range(L)]
        ws.e_xn2 = [torch.zeros(ME, d, **a) for _ in range(L)]
        ws.e_hid = [torch.zeros(ME, ff, **a) for _ in range(L)]
        dim_g = 64 if not c.no_box_trigonometric_embedding else 4
        ws.wg_eff_all = torch.zeros(L * h, dim_g, **f32)
        ws.wg_eff = [ws.wg_eff_all[l * h: (l + 1) * h] for l in range(L)]
        # all WG biases as one [L*h] view when the layout packed them without padding (h % 4 == 0)
        b0 = self._offs_w["model.encoder.layers.0.self_attn.WGs.0.bias"]
        packed = all(self._offs_w[f"model.encoder.layers.{l}.self_attn.WGs.{j}.bias"] == b0 + l * h + j for l in range(L) for j in range(h))
        ws.wg_b_all = self.flat_w[b0: b0 + L * h] if packed else None
        ws.mem = torch.zeros(ME, d, **a)
        ws.memkv = [torch.zeros(ME, 2 * d, **a) for _ in range(L)]
        ws.y = [torch.zeros(MD, d, **f32) for _ in range(3 * L + 1)]
        ws.d_yn1 = [torch.zeros(MD, d, **a) for _ in range(L)]
        ws.d_qkv = [torch.zeros(MD, 3 * d, **a) for _ in range(L)]
        ws.d_probs = [torch.zeros(R, h, T, T, **f32) for _ in range(L)]
        ws.d_att = [torch.zeros(MD, d, **a) for _ in range(L)]
        ws.d_yn2 = [torch.zeros(MD, d, **a) for _ in range(L)]
        ws.d_qc = [torch.zeros(MD, d, **a) for _ in range(L)]
        ws.c_probs = [torch.zeros(B, h, S * T, N, **f32) for _ in range(L)]
        ws.d_catt = [torch.zeros(MD, d, **a) for _ in range(L)]
        ws.d_yn3 = [torch.zeros(MD, d, **a) for _ in range(L)]
        ws.d_hid = [torch.zeros(MD, ff, **a) for _ in range(L)]
        ws.yf = torch.zeros(MD, d, **a)
        ws.logits = torch.zeros(MD, self.Vp, **f32)
        ws.dlogits = torch.zeros(MD, self.Vp, **a)
        # backward scratch
        nmax = max(self.Vp, ff, 3 * d, F)
        ws.gb = torch.zeros(max(ME, MD) * max(ff, 3 * d), **a)         # grad operand [M, N]
        # bf16: the weight-gradient GEMMs (+ their split-K reductions) run on a side stream next to the dX chain; the
        # gradient operand they read lives in a ring so that the next layers can already overwrite "their" buffer
        ws.gb_ring = [ws.gb] + [torch.zeros_like(ws.gb) for _ in range(self.wgrad_ring - 1)] if self._side_wgrad() else [ws.gb]
        ws.gb_free = [None] * len(ws.gb_ring)   # event: the side-stream consumer of ring slot i has finished
        ws.gb_next = 0
        ws.gT = torch.zeros(nmax * Mp, **a)                            # grad operand transposed [N, Mp]
        ws.xT = torch.zeros(max(ff, F, d) * Mp, **a)                   # activation transposed [K, Mp]
        ws.wT = torch.zeros(max(self.Vp * d, ff * d, F * d, 3 * d * d), **a)  # (W.m)^T [K, N]
        ws.ga = torch.zeros(max(ME, MD), max(ff, 3 * d), **f32)        # fp32 activation gradient scratch
        ws.gq = torch.zeros(max(ME, MD), 3 * d, **f32)                 # dq|dk|dv
        ws.dres = [torch.zeros(max(ME, MD), d, **f32) for _ in range(2)]  # running residual-stream gradient (ping-pong)
        ws.dmem = torch.zeros(ME, d, **f32)
        ws.dmemkv = torch.zeros(ME, 2 * d, **f32)
        ws.dbias = torch.zeros(B, h, N, N, **f32)
        ws.dwg = torch.zeros(h, ws.wg_eff[0].shape[1], **f32)
        ws.dtable = torch.zeros(V, d, **f32)
        # split-K partial products of the weight-gradient GEMMs (sc_linear_wgrad workspace): up to 8 fp32 copies of the
        # largest non-vocabulary weight, at least 2 of the generator
        ws.wgrad_ws = torch.zeros(max(8 * max(ff * d, F * d, 3 * d * d), 2 * self.Vp * d), **f32)
        self._ws[key] = ws
        return ws

    def _side_wgrad(self):
        return self.adt == torch.bfloat16 and self.wgrad_ring > 1

    def _opt_stream(self):
        if self._ost is None:
            self._ost = torch.cuda.Stream(self.dev)
        return self._ost

    def _side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream(self.dev)
        return self._side

    # ---- linear forward / backward helpers ------------------------------------------------------------------
    def _lin(self, wname, x, out, *, count=1, relu=False, residual=None, p=0.0, site=0, bias=True):
        W, S, mode, U, seed, stream = self._mask_args(wname, count)
        b = self._group(self.p, wname.replace(".weight", ".bias"), count) if bias else None
        p = p if self.training else 0.0
        if self.adt == torch.bfloat16 and self.premask:
            # bf16 training at M >= 1800 rows: W (.) m is materialised ONCE per step as a bf16 TMA operand; masking in
            # the operand prologue would redo sigmoid + Philox for every 128-row tile ("K1: where to mask")
            key = (wname, count)
            wm = self._wm.get(key)
            if wm is None:
                wm = self._wm[key] = torch.empty(W.shape, device=self.dev, dtype=torch.bfloat16)
                self._wmT[key] = torch.empty(W.shape[1], W.shape[0], device=self.dev, dtype=torch.bfloat16)
            if self._wm_step.get(key) != (self.step_id, self.training):
                if self.training:
                    # one pass over W and S: the forward operand W (.) m and the dX operand (W (.) m)^T, same mask sample
                    K.apply_mask_transposed(W, S, mode, self._wmT[key], uniforms=U, seed=seed, stream_id=stream, out=wm)
                else:
                    K.apply_mask(W, S, mode, uniforms=U, seed=seed, stream_id=stream, out=wm)
                self._wm_step[key] = (self.step_id, self.training)
            K.linear_dropout(x, wm, b, residual=residual, relu=relu, out=out, p=p, drop_seed=self._sd(1),
                             drop_stream=self._drop_stream(site), tile_n=self._fwd_tile(x.shape[0], wm.shape[0], wm.shape[1], residual is not None))
            return out
        K.linear_dropout(x, W, b, mask=S, mask_mode=mode, uniforms=U, seed=seed, stream_id=stream, residual=residual,
                         relu=relu, out=out, p=p, drop_seed=self._sd(1), drop_stream=self._drop_stream(site))
        return out

    @staticmethod
    def _fwd_tile(M, N, Kd, has_res):
        """Tile hints (1000 * stages + block_n) where the sweep over the training shapes (scripts/train_gemm_sweep.py,
        profiles/r01b_train_gemm_sweep.txt) beat the kernel's own heuristic: the fp32-residual + dropout epilogue is the long
        pole of the d x d GEMMs and likes the 8-epilogue-warp configurations."""
        if has_res and N <= 512:
            return 5128 if M >= 3000 else 6064
        if not has_res and N >= 2048 and M < 3000:
            return 3256
        return 0

    def _lin_bwd(self, ws, wname, x_saved, g, *, count=1, h=None, p=0.0, site=0, dx=None, dx_residual=None, g_ready=None, pre=None,
                 dx_next=None):
        """Backward of y = drop(act(x W^T + b)).  g: fp32 [M,N] gradient wrt the layer output (after dropout);
        h: saved post-activation output when the layer has ReLU (mask = h != 0, which also covers its dropout);
        p/site: dropout to regenerate when h is None.  g_ready: (gb, None) when the gradient is already in the
        activation dtype and needs no masking (generator).  dx (fp32 [M,K]) receives g' (W.m) [+ dx_residual]."""
        W, S, mode, U, seed, stream = self._mask_args(wname, count)
        N, Kd = W.shape
        M = x_saved.shape[0]
        Mp = K.pad8(M)
        p = p if self.training else 0.0
        bname = wname.replace(".weight", ".bias")
        gbias = self._group(self.g, bname, count) if bname in self.g else None
        # bf16: the weight-gradient GEMM reads dy [M,N] and x [M,K] as MN-major tiles - no transposed copies, no padding
        rowmajor = self.adt == torch.bfloat16 and N % 8 == 0 and Kd % 8 == 0
        gT = None if rowmajor else ws.gT[: N * Mp].view(N, Mp)
        side = rowmajor and len(ws.gb_ring) > 1
        slot = None
        if pre is not None:
            # the LayerNorm backward that produced g already wrote the masked bf16 operand and the bias gradient (_ln_bwd)
            gb, slot = pre
        elif g_ready is not None:
            gb = g_ready
            if not rowmajor:
                if Mp != M:
                    gT[:, M:].zero_()
                K.transpose(gb, gT)
            if gbias is not None:
                K.colsum(gb, gbias)
        else:
            if side:
                gb, slot = self._take_gb(ws, M, N)
            else:
                gb = ws.gb[: M * N].view(M, N)
            if not rowmajor and Mp != M:
                gT[:, M:].zero_()
            # the bias gradient (column sums) is accumulated by the same pass (flat_gw is zeroed at the start of the backward)
            K.prep_grad(g, h=h, out=gb, outT=gT, scale=(1.0 / (1.0 - p)) if (h is not None and p > 0) else 1.0,
                        p=0.0 if h is not None else p, seed=self._sd(1), stream_id=self._drop_stream(site), colsum=gbias)
        nxt_pre = None
        if dx is not None:
            wT = self._wmT.get((wname, count)) if (self.adt == torch.bfloat16 and self.premask and self.training) else None
            if wT is None:
                wT = ws.wT[: Kd * N].view(Kd, N)
                K.apply_mask_transposed(W, S, mode, wT, uniforms=U, seed=seed, stream_id=stream)
            if dx_next is not None and self.fuse_hmask and side and dx_residual is None and Kd % 8 == 0:
                # this linear's input was h = dropout(relu(previous linear)): the dX GEMM's epilogue applies that mask, stores
                # the bf16 gradient operand of the previous linear straight into a ring buffer and accumulates its bias
                # gradient (sc_linear_hmask) - no fp32 dX tensor, no separate sc_prep_grad pass
                hm, scale, next_wname = dx_next
                gb_next, slot_next = self._take_gb(ws, M, Kd)
                K.linear_hmask(gb, wT, hm, gb_next, scale=scale, colsum=self.g[next_wname.replace(".weight", ".bias")])
                nxt_pre = (gb_next, slot_next)
            else:
                K.linear(gb, wT, None, residual=dx_residual, out=dx, tile_n=5128 if (N >= 8192 and Kd <= 512) else 0)  # generator dX
        gW = self._group(self.g, wname, count)
        gS = self._group(self.gs, wname, count) if S is not None else None
        if self.fused_st:
            S, gS, mode, U = None, None, K.MASK_NONE, None  # the weight gradient stays dWm; sc_adam_clip_st applies the mask
        if side:
            # fork after the gradient operand exists (the dX GEMM above only reads it), join before the optimizer
            main, st = torch.cuda.current_stream(self.dev), self._side_stream()
            ev = torch.cuda.Event()
            ev.record(main)
            st.wait_event(ev)
            with torch.cuda.stream(st):
                if not self._diag_skip_wgrad:
                    K.linear_wgrad_rowmajor(gb, x_saved, W, S, mode, gW, gS, workspace=ws.wgrad_ws, uniforms=U, seed=seed,
                                            stream_id=stream, bypass=self.bypass)
                if slot is not None:
                    ws.gb_free[slot] = torch.cuda.Event()
                    ws.gb_free[slot].record(st)
            ws.side_used = True
            return nxt_pre
        if rowmajor:
            K.linear_wgrad_rowmajor(gb, x_saved, W, S, mode, gW, gS, workspace=ws.wgrad_ws, uniforms=U, seed=seed, stream_id=stream,
                                    bypass=self.bypass)
            return
        xT = ws.xT[: Kd * Mp].view(Kd, Mp)
        if Mp != M:
            xT[:, M:].zero_()
        K.transpose(x_saved, xT)
        K.linear_wgrad(gT, xT, W, S, mode, gW, gS, M=Mp, uniforms=U, seed=seed, stream_id=stream, bypass=self.bypass,
                       workspace=ws.wgrad_ws if self.adt == torch.bfloat16 else None)

    def _ln(self, name, x, out):
        return K.layernorm(x, self.p[name + ".a_2"], self.p[name + ".b_2"], out=out)

    def _attn_fused(self, ws):
        """attention backward writes bf16 gradient operands + bias gradients itself (tensor path: bf16, d_k = 64, ring)."""
        c = self.cfg
        return (self.fuse_attn_bwd and self.adt == torch.bfloat16 and c.d_model // c.num_heads == 64 and len(ws.gb_ring) > 1
                and ws.N <= 128 and ws.T <= 128)

    def _take_gb(self, ws, M, N):
        """Next gradient-operand buffer of the ring (waits for the side-stream consumer that used it last)."""
        slot = ws.gb_next
        ws.gb_next = (slot + 1) % len(ws.gb_ring)
        if ws.gb_free[slot] is not None:
            torch.cuda.current_stream(self.dev).wait_event(ws.gb_free[slot])
        return ws.gb_ring[slot][: M * N].view(M, N), slot

    def _ln_bwd(self, name, x, dy, dx, dres=None, nxt=None, ws=None):
        """Backward of LayerNorm `name`.  ``nxt = (weight name, dropout p, site)``: the linear whose output gradient is dx
        comes next in the backward chain - its bf16 gradient operand (dropout mask regenerated) and bias gradient are
        produced by the same kernel; returns ``pre`` for ``_lin_bwd`` (None when not fused)."""
        d = x.shape[1]
        if nxt is None or self.adt != torch.bfloat16 or d != 512 or len(ws.gb_ring) < 2:
            K.layernorm_bwd(x, self.p[name + ".a_2"], dy, dx, self.g[name + ".a_2"], self.g[name + ".b_2"], dres=dres)
            return None
        wname, p, site = nxt
        gb, slot = self._take_gb(ws, x.shape[0], d)
        gbias = self.g[wname.replace(".weight", ".bias")]
        K.layernorm_bwd(x, self.p[name + ".a_2"], dy, dx, self.g[name + ".a_2"], self.g[name + ".b_2"], dres=dres, next_gb=gb,
                        next_colsum=gbias, next_p=p if self.training else 0.0, seed=self._sd(1), stream_id=self._drop_stream(site))
        return gb, slot

    # ---------------------------------------------------------------------------------------------------------
    def load_batch(self, ws, att_feats, boxes, seqs, masks, att_masks=None, global_tokens=None):
        c = self.cfg
        B, N, S, T = ws.B, ws.N, ws.S, ws.T
        ws.att_in.copy_(att_feats.reshape(B * N, -1), non_blocking=True)
        ws.boxes.copy_(boxes, non_blocking=True)
        if att_masks is not None:
            ws.att_mask.copy_(att_masks.float(), non_blocking=True)
        seqs = seqs.to(self.dev, non_blocking=True)
        masks = masks.to(self.dev, non_blocking=True).float()
        tok = seqs[:, :T].contiguous()
        ws.tokens.copy_(tok.reshape(-1).int())
        ws.targets.copy_(seqs[:, 1: T + 1].reshape(-1).int())
        ws.tok_w.copy_(masks[:, 1: T + 1].reshape(-1))
        ws.key_valid.copy_((tok != c.pad_token_id).float())
        denom = masks[:, 1: T + 1].sum() if global_tokens is None else global_tokens
        ws.inv_norm.copy_((1.0 / denom).reshape(1))

    def forward(self, ws):
        """Teacher-forcing forward; leaves logits in ws.logits and every activation the backward needs."""
        c = self.cfg
        B, N, S, T, ME, MD, R = ws.B, ws.N, ws.S, ws.T, ws.ME, ws.MD, ws.R
        d, h, L = c.d_model, c.num_heads, c.num_layers
        dk = d // h
        pd = self.p_drop if self.training else 0.0
        trig = not c.no_box_trigonometric_embedding
        if ws.att_a is not ws.att_in:
            K.cast_bf16(ws.att_in, out=ws.att_a)
        if self.adt == torch.bfloat16 and self.premask and self.training:
            self._premask_all()
        self._lin("att_embed.0.weight", ws.att_a, ws.xe[0], relu=True, p=self.p_src, site=1)
        if ws.att_mask is not None:
            K.mask_rows(ws.xe[0], ws.att_mask.view(-1))
        # effective (masked) WG weights of every layer, then the log-geometry bias of all layers and heads in ONE pass over
        # the box pairs (the sin/cos embedding only depends on the boxes; the reference rebuilds it in each layer)
        for l in range(L):
            Wg, Sg, mode, U, seed, stream = self._mask_args(f"model.encoder.layers.{l}.self_attn.WGs.0.weight", h)
            if Sg is None:
                ws.wg_eff[l].copy_(Wg)
            else:
                K.apply_mask(Wg, Sg, mode, uniforms=U, seed=seed, stream_id=stream, out=ws.wg_eff[l])
        if ws.wg_b_all is not None:
            K.box_bias_all(ws.boxes, ws.wg_eff_all, ws.wg_b_all, ws.e_bias_all, B=B, N=N, layers=L, h=h, trig=trig)
        for l in range(L):
            p = f"model.encoder.layers.{l}"
            x0, x1, x2 = ws.xe[2 * l], ws.xe[2 * l + 1], ws.xe[2 * l + 2]
            self._ln(f"{p}.sublayer.0.norm", x0, ws.e_xn1[l])
            self._lin(f"{p}.self_attn.linears.0.weight", ws.e_xn1[l], ws.e_qkv[l], count=3)
            if ws.wg_b_all is None:
                K.box_bias_fwd(ws.boxes, ws.wg_eff[l], self._group(self.p, f"{p}.self_attn.WGs.0.bias", h), ws.e_bias[l], B=B, N=N,
                               h=h, trig=trig)
            q = ws.e_qkv[l]
            K.attention_fwd(q[:, 0:], q[:, d:], q[:, 2 * d:], ws.e_att[l], ws.e_probs[l], G=B, Tq=N, Tk=N, h=h, dk=dk, ldq=3 * d,
                            ldk=3 * d, ldv=3 * d, ldo=d, key_valid=ws.att_mask, bias=ws.e_bias[l], p=pd, seed=self._sd(2),
                            stream_id=self._drop_stream(10 + l))
            self._lin(f"{p}.self_attn.linears.3.weight", ws.e_att[l], x1, residual=x0, p=pd, site=20 + l)
            self._ln(f"{p}.sublayer.1.norm", x1, ws.e_xn2[l])
            self._lin(f"{p}.feed_forward.w_1.weight", ws.e_xn2[l], ws.e_hid[l], relu=True, p=pd, site=30 + l)
            self._lin(f"{p}.feed_forward.w_2.weight", ws.e_hid[l], x2, residual=x1, p=pd, site=40 + l)
        self._ln("model.encoder.norm", ws.xe[2 * L], ws.mem)
        for l in range(L):
            self._lin(f"model.decoder.layers.{l}.src_attn.linears.1.weight", ws.mem, ws.memkv[l], count=2)
        # ---- decoder ----
        W, S_, mode, U, seed, stream = self._mask_args("model.tgt_embed.0.lut.weight")
        K.embed_pe(ws.tokens, W, self.pe, T=T, pos0=0, mask=S_, mask_mode=mode, uniforms=U, seed=seed, stream_id=stream, out=ws.y[0])
        if pd > 0:
            K.prep_grad(ws.y[0], out=ws.y[0], p=pd, seed=self._sd(1), stream_id=self._drop_stream(2))
        for l in range(L):
            p = f"model.decoder.layers.{l}"
            y0, y1, y2, y3 = ws.y[3 * l], ws.y[3 * l + 1], ws.y[3 * l + 2], ws.y[3 * l + 3]
            self._ln(f"{p}.sublayer.0.norm", y0, ws.d_yn1[l])
            self._lin(f"{p}.self_attn.linears.0.weight", ws.d_yn1[l], ws.d_qkv[l], count=3)
            q = ws.d_qkv[l]
            K.attention_fwd(q[:, 0:], q[:, d:], q[:, 2 * d:], ws.d_att[l], ws.d_probs[l], G=R, Tq=T, Tk=T, h=h, dk=dk, ldq=3 * d,
                            ldk=3 * d, ldv=3 * d, ldo=d, key_valid=ws.key_valid, causal_T=T, p=pd, seed=self._sd(2),
                            stream_id=self._drop_stream(50 + l))
            self._lin(f"{p}.self_attn.linears.3.weight", ws.d_att[l], y1, residual=y0, p=pd, site=60 + l)
            self._ln(f"{p}.sublayer.1.norm", y1, ws.d_yn2[l])
            self._lin(f"{p}.src_attn.linears.0.weight", ws.d_yn2[l], ws.d_qc[l])
            kv = ws.memkv[l]
            K.attention_fwd(ws.d_qc[l], kv[:, 0:], kv[:, d:], ws.d_catt[l], ws.c_probs[l], G=B, Tq=S * T, Tk=N, h=h, dk=dk, ldq=d,
                            ldk=2 * d, ldv=2 * d, ldo=d, key_valid=ws.att_mask, p=pd, seed=self._sd(2),
                            stream_id=self._drop_stream(70 + l))
            self._lin(f"{p}.src_attn.linears.3.weight", ws.d_catt[l], y2, residual=y1, p=pd, site=80 + l)
            self._ln(f"{p}.sublayer.2.norm", y2, ws.d_yn3[l])
            self._lin(f"{p}.feed_forward.w_1.weight", ws.d_yn3[l], ws.d_hid[l], relu=True, p=pd, site=90 + l)
            self._lin(f"{p}.feed_forward.w_2.weight", ws.d_hid[l], y3, residual=y2, p=pd, site=100 + l)
        self._ln("model.decoder.norm", ws.y[3 * L], ws.yf)
        self._lin("model.generator.proj.weight", ws.yf, ws.logits)
        return ws.logits

    def loss_and_backward(self, ws):
        """LanguageModelCriterion + full backward; gradients land in flat_gw / flat_gs (overwritten, not accumulated)."""
        for phase in range(self.N_PHASES):
            self.backward_phase(ws, phase)
        return ws.loss_sum

    # The backward in three phases whose gradients are final when the phase ends, so that a data-parallel run can start
    # the all-reduce of a finished bucket while the next phase computes (grad_buckets):
    #   0: loss, generator, decoder norm, decoder layers L-1 .. L/2     1: decoder layers L/2-1 .. 0, embedding
    #   2: encoder norm, encoder layers L-1 .. L/2                      3: encoder layers L/2-1 .. 0, att_embed
    N_PHASES = 4

    def grad_buckets(self, phase):
        """[(flat gradient buffer, start, end)] of the parameter ranges whose gradients are final after `phase`."""
        L = self.cfg.num_layers
        half = L // 2
        def rng(offs, keys, first, last):
            # [offset of `first`, offset of `last`) in a layout; names that are not in the layout (norms / biases in the
            # logit layout) move forward to the next name that is
            def at(name):
                if name is None:
                    return None
                i = self.names.index(name)
                while i < len(self.names) and self.names[i] not in offs:
                    i += 1
                return offs[self.names[i]] if i < len(self.names) else None
            return at(first), at(last)
        dec0, dech = "model.decoder.layers.0.self_attn.linears.0.weight", f"model.decoder.layers.{half}.self_attn.linears.0.weight"
        lut, gen = "model.tgt_embed.0.lut.weight", "model.generator.proj.weight"
        ench = f"model.encoder.layers.{half}.self_attn.linears.0.weight"
        spans = {0: [(dech, lut), (gen, None)], 1: [(dec0, dech), (lut, gen)], 2: [(ench, dec0)], 3: [(self.names[0], ench)]}[phase]
        out = []
        for flat, offs in ((self.flat_gw, self._offs_w), (self.flat_gs, self._offs_s) if (self.masked and not self.fused_st) else (None, None)):
            if flat is None:
                continue
            for first, last in spans:
                a, b = rng(offs, None, first, last)
                a = 0 if a is None else a
                b = flat.numel() if b is None else b
                if b > a:
                    out.append((flat, a, b))
        return out

    def backward_phase(self, ws, phase):
        c = self.cfg
        B, N, S, T, ME, MD, R = ws.B, ws.N, ws.S, ws.T, ws.ME, ws.MD, ws.R
        d, h, L, ff = c.d_model, c.num_heads, c.num_layers, c.dim_feedforward
        dk = d // h
        pd = self.p_drop if self.training else 0.0
        trig = not c.no_box_trigonometric_embedding
        ws.gb_free = [None] * len(ws.gb_ring)
        ws.gb_next = 0
        ws.side_used = False
        half = L // 2
        if phase == 0:
            self._materialized = False
            # norm and bias gradients accumulate through atomics: one memset of the flat buffer (weights are overwritten)
            self.flat_gw.zero_()
            ws.loss_sum.zero_()
            K.logsoftmax_nll(ws.logits, ws.targets, ws.tok_w, ws.inv_norm, ws.loss_sum, ws.dlogits)
            ga_d = ws.ga.view(-1)[: MD * d].view(MD, d)
            self._lin_bwd(ws, "model.generator.proj.weight", ws.yf, None, g_ready=ws.dlogits, dx=ga_d)
            ws.cur = 0
            pre = self._ln_bwd("model.decoder.norm", ws.y[3 * L], ga_d, ws.dres[0][:MD], ws=ws,
                               nxt=(f"model.decoder.layers.{L - 1}.feed_forward.w_2.weight", pd, 100 + L - 1))
            ws.dmem.zero_()
        if phase >= 2:
            self._backward_encoder(ws, phase - 2)
            return
        if phase != 0:
            pre = None  # (a phase starts with a plain gradient preparation: the ring was reset)
        first = half if phase == 0 else 0  # last layer this phase visits
        cur, nxt = ws.dres[ws.cur][:MD], ws.dres[1 - ws.cur][:MD]
        for l in (reversed(range(half, L)) if phase == 0 else reversed(range(half))):
            p = f"model.decoder.layers.{l}"
            y0, y1, y2 = ws.y[3 * l], ws.y[3 * l + 1], ws.y[3 * l + 2]
            ga_ff = ws.ga.view(-1)[: MD * ff].view(MD, ff)
            # feed-forward sublayer
            pre1 = self._lin_bwd(ws, f"{p}.feed_forward.w_2.weight", ws.d_hid[l], cur, p=pd, site=100 + l, dx=ga_ff, pre=pre,
                                 dx_next=(ws.d_hid[l], 1.0 / (1.0 - pd) if pd > 0 else 1.0, f"{p}.feed_forward.w_1.weight"))
            ga_dd = ws.gq.view(-1)[: MD * d].view(MD, d)
            self._lin_bwd(ws, f"{p}.feed_forward.w_1.weight", ws.d_yn3[l], ga_ff, h=ws.d_hid[l], p=pd, dx=ga_dd, pre=pre1)
            pre = self._ln_bwd(f"{p}.sublayer.2.norm", y2, ga_dd, nxt, dres=cur, ws=ws, nxt=(f"{p}.src_attn.linears.3.weight", pd, 80 + l))
            cur, nxt = nxt, cur
            # cross-attention sublayer
            ga_c = ws.ga.view(-1)[: MD * d].view(MD, d)
            self._lin_bwd(ws, f"{p}.src_attn.linears.3.weight", ws.d_catt[l], cur, p=pd, site=80 + l, dx=ga_c, pre=pre)
            kv = ws.memkv[l]
            dqc = ws.gq.view(-1)[: MD * d].view(MD, d)
            pre_q = pre_kv = None
            if self._attn_fused(ws):
                gbq, sq = self._take_gb(ws, MD, d)
                gbkv, skv = self._take_gb(ws, ME, 2 * d)
                bkv = self._group(self.g, f"{p}.src_attn.linears.1.bias", 2)
                K.attention_bwd_bf16out(ws.d_qc[l], kv[:, 0:], kv[:, d:], ws.c_probs[l], ga_c, gbq, gbkv[:, 0:], gbkv[:, d:], G=B,
                                        Tq=S * T, Tk=N, h=h, dk=dk, ldq=d, ldk=2 * d, ldv=2 * d, ldd=d, ldgq=d, ldgk=2 * d, ldgv=2 * d,
                                        bq=self.g[f"{p}.src_attn.linears.0.bias"], bk=bkv[:d], bv=bkv[d:], p=pd, seed=self._sd(2),
                                        stream_id=self._drop_stream(70 + l))
                pre_q, pre_kv = (gbq, sq), (gbkv, skv)
            else:
                K.attention_bwd(ws.d_qc[l], kv[:, 0:], kv[:, d:], ws.c_probs[l], ga_c, dqc, ws.dmemkv[:, 0:], ws.dmemkv[:, d:],
                                dtype=self.adt, G=B, Tq=S * T, Tk=N, h=h, dk=dk, ldq=d, ldk=2 * d, ldv=2 * d, ldd=d, ldgq=d, ldgk=2 * d,
                                ldgv=2 * d, p=pd, seed=self._sd(2), stream_id=self._drop_stream(70 + l))
            ga_c2 = ws.ga.view(-1)[: MD * d].view(MD, d)
            self._lin_bwd(ws, f"{p}.src_attn.linears.0.weight", ws.d_yn2[l], dqc, dx=ga_c2, pre=pre_q)
            pre = self._ln_bwd(f"{p}.sublayer.1.norm", y1, ga_c2, nxt, dres=cur, ws=ws, nxt=(f"{p}.self_attn.linears.3.weight", pd, 60 + l))
            cur, nxt = nxt, cur
            self._lin_bwd(ws, f"{p}.src_attn.linears.1.weight", ws.mem, ws.dmemkv, count=2, dx=ws.dmem, dx_residual=ws.dmem, pre=pre_kv)
            # self-attention sublayer
            ga_s = ws.ga.view(-1)[: MD * d].view(MD, d)
            self._lin_bwd(ws, f"{p}.self_attn.linears.3.weight", ws.d_att[l], cur, p=pd, site=60 + l, dx=ga_s, pre=pre)
            q = ws.d_qkv[l]
            gq = ws.gq[:MD]
            pre_qkv = None
            if self._attn_fused(ws):
                gb3, s3 = self._take_gb(ws, MD, 3 * d)
                b3 = self._group(self.g, f"{p}.self_attn.linears.0.bias", 3)
                K.attention_bwd_bf16out(q[:, 0:], q[:, d:], q[:, 2 * d:], ws.d_probs[l], ga_s, gb3[:, 0:], gb3[:, d:], gb3[:, 2 * d:], G=R,
                                        Tq=T, Tk=T, h=h, dk=dk, ldq=3 * d, ldk=3 * d, ldv=3 * d, ldd=d, ldgq=3 * d, ldgk=3 * d,
                                        ldgv=3 * d, bq=b3[:d], bk=b3[d: 2 * d], bv=b3[2 * d:], p=pd, seed=self._sd(2),
                                        stream_id=self._drop_stream(50 + l))
                pre_qkv = (gb3, s3)
            else:
                K.attention_bwd(q[:, 0:], q[:, d:], q[:, 2 * d:], ws.d_probs[l], ga_s, gq[:, 0:], gq[:, d:], gq[:, 2 * d:], dtype=self.adt,
                                G=R, Tq=T, Tk=T, h=h, dk=dk, ldq=3 * d, ldk=3 * d, ldv=3 * d, ldd=d, ldgq=3 * d, ldgk=3 * d, ldgv=3 * d,
                                p=pd, seed=self._sd(2), stream_id=self._drop_stream(50 + l))
            ga_s2 = ws.ga.view(-1)[: MD * d].view(MD, d)
            self._lin_bwd(ws, f"{p}.self_attn.linears.0.weight", ws.d_yn1[l], gq, count=3, dx=ga_s2, pre=pre_qkv)
            pre = self._ln_bwd(f"{p}.sublayer.0.norm", y0, ga_s2, nxt, dres=cur, ws=ws,
                               nxt=(f"model.decoder.layers.{l - 1}.feed_forward.w_2.weight", pd, 100 + l - 1) if l > first else None)
            cur, nxt = nxt, cur
        ws.cur = 0 if cur.data_ptr() == ws.dres[0].data_ptr() else 1
        if phase == 0:
            self._join_side(ws)
            return
        # embedding: its dropout mask is regenerated from the forward's Philox stream (an exact zero in the saved output is
        # not proof of a dropped element: sin(0) = 0 in the positional encoding, masked-out table entries)
        W, S_, mode, U, seed, stream = self._mask_args("model.tgt_embed.0.lut.weight")
        emb_g = ws.ga.view(-1)[: MD * d].view(MD, d)
        if pd > 0:
            K.prep_grad(cur, out=emb_g, p=pd, seed=self._sd(1), stream_id=self._drop_stream(2))
        else:
            emb_g = cur
        if self.fused_st:
            # (flat_gw was zeroed at the start of the backward: the scatter-add lands directly in the table's gradient view)
            K.embedding_bwd(ws.tokens, emb_g, self.g["model.tgt_embed.0.lut.weight"], math.sqrt(d))
        else:
            ws.dtable.zero_()
            K.embedding_bwd(ws.tokens, emb_g, ws.dtable, math.sqrt(d))
            K.mask_grad(ws.dtable, W, S_, mode, self.g["model.tgt_embed.0.lut.weight"],
                        self.gs.get("model.tgt_embed.0.lut.weight"), uniforms=U, seed=seed, stream_id=stream, bypass=self.bypass)
        self._join_side(ws)

    def _join_side(self, ws):
        if ws.side_used:
            torch.cuda.current_stream(self.dev).wait_stream(self._side_stream())
            ws.side_used = False

    def _backward_encoder(self, ws, part):
        """part 0: encoder norm + layers L-1 .. L/2 ; part 1: layers L/2-1 .. 0 + att_embed."""
        c = self.cfg
        B, N, S, T, ME, MD, R = ws.B, ws.N, ws.S, ws.T, ws.ME, ws.MD, ws.R
        d, h, L, ff = c.d_model, c.num_heads, c.num_layers, c.dim_feedforward
        dk = d // h
        pd = self.p_drop if self.training else 0.0
        trig = not c.no_box_trigonometric_embedding
        half = L // 2
        pre = None
        if part == 0:
            ws.cur = 0
            pre = self._ln_bwd("model.encoder.norm", ws.xe[2 * L], ws.dmem, ws.dres[0][:ME], ws=ws,
                               nxt=(f"model.encoder.layers.{L - 1}.feed_forward.w_2.weight", pd, 40 + L - 1))
        first = half if part == 0 else 0
        cur, nxt = ws.dres[ws.cur][:ME], ws.dres[1 - ws.cur][:ME]
        for l in (reversed(range(half, L)) if part == 0 else reversed(range(half))):
            p = f"model.encoder.layers.{l}"
            x0, x1 = ws.xe[2 * l], ws.xe[2 * l + 1]
            ga_ff = ws.ga.view(-1)[: ME * ff].view(ME, ff)
            pre1 = self._lin_bwd(ws, f"{p}.feed_forward.w_2.weight", ws.e_hid[l], cur, p=pd, site=40 + l, dx=ga_ff, pre=pre,
                                 dx_next=(ws.e_hid[l], 1.0 / (1.0 - pd) if pd > 0 else 1.0, f"{p}.feed_forward.w_1.weight"))
            ga_dd = ws.gq.view(-1)[: ME * d].view(ME, d)
            self._lin_bwd(ws, f"{p}.feed_forward.w_1.weight", ws.e_xn2[l], ga_ff, h=ws.e_hid[l], p=pd, dx=ga_dd, pre=pre1)
            pre = self._ln_bwd(f"{p}.sublayer.1.norm", x1, ga_dd, nxt, dres=cur, ws=ws, nxt=(f"{p}.self_attn.linears.3.weight", pd, 20 + l))
            cur, nxt = nxt, cur
            ga_a = ws.ga.view(-1)[: ME * d].view(ME, d)
            self._lin_bwd(ws, f"{p}.self_attn.linears.3.weight", ws.e_att[l], cur, p=pd, site=20 + l, dx=ga_a, pre=pre)
            q = ws.e_qkv[l]
            gq = ws.gq[:ME]
            pre_qkv = None
            if self._attn_fused(ws):
                gb3, s3 = self._take_gb(ws, ME, 3 * d)
                b3 = self._group(self.g, f"{p}.self_attn.linears.0.bias", 3)
                K.attention_bwd_bf16out(q[:, 0:], q[:, d:], q[:, 2 * d:], ws.e_probs[l], ga_a, gb3[:, 0:], gb3[:, d:], gb3[:, 2 * d:], G=B,
                                        Tq=N, Tk=N, h=h, dk=dk, ldq=3 * d, ldk=3 * d, ldv=3 * d, ldd=d, ldgq=3 * d, ldgk=3 * d,
                                        ldgv=3 * d, bq=b3[:d], bk=b3[d: 2 * d], bv=b3[2 * d:], dbias=ws.dbias, p=pd, seed=self._sd(2),
                                        stream_id=self._drop_stream(10 + l))
                pre_qkv = (gb3, s3)
            else:
                K.attention_bwd(q[:, 0:], q[:, d:], q[:, 2 * d:], ws.e_probs[l], ga_a, gq[:, 0:], gq[:, d:], gq[:, 2 * d:], dtype=self.adt,
                                G=B, Tq=N, Tk=N, h=h, dk=dk, ldq=3 * d, ldk=3 * d, ldv=3 * d, ldd=d, ldgq=3 * d, ldgk=3 * d, ldgv=3 * d,
                                dbias=ws.dbias, p=pd, seed=self._sd(2), stream_id=self._drop_stream(10 + l))
            gb_wg = self._group(self.g, f"{p}.self_attn.WGs.0.bias", h)
            gb_wg.zero_()
            if self.fused_st:
                gw_wg = self._group(self.g, f"{p}.self_attn.WGs.0.weight", h)
                gw_wg.zero_()
                K.box_bias_bwd(ws.boxes, ws.e_bias[l], ws.dbias, gw_wg, gb_wg, B=B, N=N, h=h, trig=trig)
            else:
                ws.dwg.zero_()
                K.box_bias_bwd(ws.boxes, ws.e_bias[l], ws.dbias, ws.dwg, gb_wg, B=B, N=N, h=h, trig=trig)
                Wg, Sg, mode, U, seed, stream = self._mask_args(f"{p}.self_attn.WGs.0.weight", h)
                K.mask_grad(ws.dwg, Wg, Sg, mode, self._group(self.g, f"{p}.self_attn.WGs.0.weight", h),
                            self._group(self.gs, f"{p}.self_attn.WGs.0.weight", h) if Sg is not None else None, uniforms=U, seed=seed,
                            stream_id=stream, bypass=self.bypass)
            ga_a2 = ws.ga.view(-1)[: ME * d].view(ME, d)
            self._lin_bwd(ws, f"{p}.self_attn.linears.0.weight", ws.e_xn1[l], gq, count=3, dx=ga_a2, pre=pre_qkv)
            pre = self._ln_bwd(f"{p}.sublayer.0.norm", x0, ga_a2, nxt, dres=cur, ws=ws,
                               nxt=(f"model.encoder.layers.{l - 1}.feed_forward.w_2.weight", pd, 40 + l - 1) if l > first else None)
            cur, nxt = nxt, cur
        ws.cur = 0 if cur.data_ptr() == ws.dres[0].data_ptr() else 1
        if part == 0:
            self._join_side(ws)
            return
        if ws.att_mask is not None:
            K.mask_rows(cur, ws.att_mask.view(-1))
        self._lin_bwd(ws, "att_embed.0.weight", ws.att_a, cur, h=ws.xe[0], p=self.p_src if self.training else 0.0)
        self._join_side(ws)

    # ---------------------------------------------------------------------------------------------------------
    def optimizer_step(self, *, lr, mask_lr=100.0, clip=0.1, betas=(0.9, 0.98), eps=1e-9, mask_eps=1e-2, weight_decay=0.0,
                       grad_scale=1.0, sparsity_target=None, sparsity_weight=0.0, current_step=0, max_step=1, _dyn=False, part="all"):
        """clip_gradient(0.1) + Adam for the two parameter groups of train_n_prune_transformer.py:67-82, with the
        sparsity-loss gradient (prune.py:228-269) folded into the mask-logit update.  ``_dyn``: graph mode - lr, the Adam
        bias corrections and the sparsity scale come from ``self._dyn_dev`` (``_upload_step``), no host bookkeeping.
        ``part``: "all", or the two halves of an overlapped update - "hi" = decoder + embedding + generator (their gradients are
        final after backward phase 1; also computes the sparsity coefficient, which must see the logits before ANY update),
        "lo" = att_embed + encoder (after the last phase; also re-pins the padding logits)."""
        if not _dyn and part != "lo":
            self.opt_step += 1
        dyn = self._dyn_dev if _dyn else None
        if self.fused_st and not self._materialized:
            return self._optimizer_step_st(lr=lr, mask_lr=mask_lr, clip=clip, betas=betas, eps=eps, mask_eps=mask_eps,
                                           weight_decay=weight_decay, grad_scale=grad_scale, sparsity_target=sparsity_target,
                                           sparsity_weight=sparsity_weight, current_step=current_step, max_step=max_step, dyn=dyn, part=part)
        cut_w = self._offs_w["model.decoder.layers.0.self_attn.linears.0.weight"]
        lo_w, hi_w = {"all": (0, self.flat_w.numel()), "hi": (cut_w, self.flat_w.numel()), "lo": (0, cut_w)}[part]
        K.adam_clip(self.flat_w[lo_w:hi_w], self.flat_gw[lo_w:hi_w], self.m_w[lo_w:hi_w], self.v_w[lo_w:hi_w], lr=lr, betas=betas, eps=eps,
                    weight_decay=weight_decay, clip=clip, grad_scale=grad_scale, step=max(1, self.opt_step),
                    dyn=dyn[0:3] if _dyn else None)
        if self.masked and self.mask_type == "supermask":
            coeff = None
            if sparsity_target is not None and sparsity_weight:
                if part != "lo":
                    anneal = (1.0 + math.cos(min(1.0, current_step / max_step) * math.pi)) / 2.0
                    self.sp_count.zero_()
                    K.lib.call("sc_mask_count", K.lib.ptr(self.flat_s), self.flat_s.numel(), K.lib.ptr(self.sp_count), K.lib.stream())
                    K.sparsity_coeff(self.sp_count, self.n_logits, sparsity_target, sparsity_weight * (1.0 - anneal), self.sp_out,
                                     scale_dev=dyn[6:7] if _dyn else None)
                coeff = self.sp_out[1:2]
            cut_s = self._offs_s["model.decoder.layers.0.self_attn.linears.0.weight"]
            lo_s, hi_s = {"all": (0, self.flat_s.numel()), "hi": (cut_s, self.flat_s.numel()), "lo": (0, cut_s)}[part]
            K.adam_clip(self.flat_s[lo_s:hi_s], self.flat_gs[lo_s:hi_s], self.m_s[lo_s:hi_s], self.v_s[lo_s:hi_s], lr=mask_lr, betas=betas,
                        eps=mask_eps, weight_decay=0.0, clip=clip, grad_scale=grad_scale, step=max(1, self.opt_step),
                        sigmoid_grad_coeff=coeff, dyn=dyn[3:6] if _dyn else None)
            if part != "hi" and self._s_pad_idx.numel():
                self.flat_s.index_fill_(0, self._s_pad_idx, -1.0)

    def _mask_groups(self):
        """(first weight name, fused count) of every masked tensor exactly as the forward samples it (one Philox stream and one
        element numbering per group): the linears of _premask_keys, the embedding table, the h geometry heads of each layer."""
        h = self.cfg.num_heads
        groups = list(self._premask_keys()) + [("model.tgt_embed.0.lut.weight", 1)]
        groups += [(f"model.encoder.layers.{l}.self_attn.WGs.0.weight", h) for l in range(self.cfg.num_layers)]
        return groups

    def _st_table(self, part):
        """Descriptor table of sc_adam_clip_st for "all" / "hi" (decoder, embedding, generator) / "lo" (att_embed, encoder)."""
        if part not in self._st_desc:
            first_of = {}
            member = set()
            if self.masked:
                for wname, count in self._mask_groups():
                    first_of[wname] = count
                    i = self.names.index(wname)
                    member.update(self.names[i + 1: i + count])  # (the fused members are consecutive names: [q;k;v], [k;v], WG heads)
            cut = self.names.index("model.decoder.layers.0.self_attn.linears.0.weight")
            names = {"all": self.names, "hi": self.names[cut:], "lo": self.names[:cut]}[part]
            segs = []
            for k in names:
                if k in member:
                    continue
                if k in first_of:
                    n = self._group(self.p, k, first_of[k]).numel()
                    segs.append((self._offs_w[k], self._offs_s[k], n, self.stream_of[k]))
                else:
                    n = self._group(self.p, k, 1).numel() if k in self._pad_rows else self.p[k].numel()
                    segs.append((self._offs_w[k], -1, n, 0))
            self._st_desc[part] = K.st_descriptors(segs, self.dev)
        return self._st_desc[part]

    def _optimizer_step_st(self, *, lr, mask_lr, clip, betas, eps, mask_eps, weight_decay, grad_scale, sparsity_target, sparsity_weight,
                           current_step, max_step, dyn, part):
        supermask = bool(self.masked) and self.mask_type == "supermask"
        coeff = None
        if supermask and sparsity_target is not None and sparsity_weight:
            if part != "lo":
                anneal = (1.0 + math.cos(min(1.0, current_step / max_step) * math.pi)) / 2.0
                self.sp_count.zero_()
                K.lib.call("sc_mask_count", K.lib.ptr(self.flat_s), self.flat_s.numel(), K.lib.ptr(self.sp_count), K.lib.stream())
                K.sparsity_coeff(self.sp_count, self.n_logits, sparsity_target, sparsity_weight * (1.0 - anneal), self.sp_out,
                                 scale_dev=dyn[6:7] if dyn is not None else None)
            coeff = self.sp_out[1:2]
        desc, blocks = self._st_table(part)
        mode = K.MASK_NONE  # the mask sample of the step being applied (train-mode forward)
        if self.masked:
            mode = (K.MASK_UNIFORM if self.flat_u is not None else K.MASK_BERNOULLI) if self.mask_type == "supermask" else K.MASK_RAW
        K.adam_clip_st(desc, blocks, self.flat_w, self.flat_gw, self.m_w, self.v_w, self.flat_s, self.m_s, self.v_s, uniforms=self.flat_u,
                       mask_mode=mode, bypass=self.bypass, update_logits=supermask, seed=self._sd(0), stream_base=self._step_base(),
                       lr=lr, eps=eps, weight_decay=weight_decay, mask_lr=mask_lr, mask_eps=mask_eps, betas=betas, clip=clip,
                       grad_scale=grad_scale, step=max(1, self.opt_step), sigmoid_grad_coeff=coeff, dyn=dyn)
        if supermask and part != "hi" and self._s_pad_idx.numel():
            self.flat_s.index_fill_(0, self._s_pad_idx, -1.0)

    def materialize_grads(self):
        """Reference ``.grad`` semantics for inspection: turns the dWm left by the backward into dW (in place, ``self.g``) and dS
        (``self.gs``) with the step's mask sample; a following ``optimizer_step`` then runs the plain two-group update on them."""
        if not self.fused_st or self._materialized or not self.masked:
            self._materialized = True
            return
        for wname, count in self._mask_groups():
            W, S, mode, U, seed, stream = self._mask_args(wname, count)
            gW = self._group(self.g, wname, count)
            K.mask_grad(gW, W, S, mode, gW, self._group(self.gs, wname, count), uniforms=U, seed=seed, stream_id=stream, bypass=self.bypass)
        self._materialized = True

    def _sparsity_coeff(self, *, sparsity_target=None, sparsity_weight=0.0, current_step=0, max_step=1, **_):
        """Device scalar d(sparsity loss)/d(nnz) from the CURRENT logits (must run before any logit is updated)."""
        if not (self.masked and self.mask_type == "supermask") or sparsity_target is None or not sparsity_weight:
            return None
        anneal = (1.0 + math.cos(min(1.0, current_step / max_step) * math.pi)) / 2.0
        self.sp_count.zero_()
        K.lib.call("sc_mask_count", K.lib.ptr(self.flat_s), self.flat_s.numel(), K.lib.ptr(self.sp_count), K.lib.stream())
        K.sparsity_coeff(self.sp_count, self.n_logits, sparsity_target, sparsity_weight * (1.0 - anneal), self.sp_out)
        return self.sp_out[1:2]

    def _sharded_step(self, ws, exchange, opt, lr):
        """Data-parallel step with the optimizer sharded over the ranks (distributed.ShardedExchange): after each backward
        phase the finished buckets go through reduce-scatter -> Adam on the owned shard -> all-gather on the exchange's
        stream while the next phase computes."""
        assert not self.fused_st, "ShardedExchange cuts tensors into rank shards: build the trainer with fused_st=False"
        o = dict(mask_lr=100.0, clip=0.1, betas=(0.9, 0.98), eps=1e-9, mask_eps=1e-2, weight_decay=0.0, grad_scale=1.0)
        o.update({k: v for k, v in opt.items() if k in o})
        coeff = self._sparsity_coeff(**opt)
        self.opt_step += 1
        step = max(1, self.opt_step)

        def upd_w(lo, hi):
            K.adam_clip(self.flat_w[lo:hi], self.flat_gw[lo:hi], self.m_w[lo:hi], self.v_w[lo:hi], lr=lr, betas=o["betas"], eps=o["eps"],
                        weight_decay=o["weight_decay"], clip=o["clip"], grad_scale=o["grad_scale"], step=step)

        def upd_s(lo, hi):
            K.adam_clip(self.flat_s[lo:hi], self.flat_gs[lo:hi], self.m_s[lo:hi], self.v_s[lo:hi], lr=o["mask_lr"], betas=o["betas"],
                        eps=o["mask_eps"], weight_decay=0.0, clip=o["clip"], grad_scale=o["grad_scale"], step=step,
                        sigmoid_grad_coeff=coeff)

        exchange.finish()  # (the coefficient above must be ordered before the first bucket's update: same stream chain)
        for phase, part in enumerate(("fwd_p0", "p1", "p2", "p3")):
            if self.use_graph:
                self._run_graphed(ws, part, opt)
            else:
                if phase == 0:
                    self.forward(ws)
                self.backward_phase(ws, phase)
            for flat, a, b in self.grad_buckets(phase):
                if flat is self.flat_gw:
                    exchange.bucket(self.flat_gw, self.flat_w, a, b, upd_w)
                elif self.mask_type == "supermask":
                    exchange.bucket(self.flat_gs, self.flat_s, a, b, upd_s)
        exchange.finish()
        if self.masked and self.mask_type == "supermask" and self._s_pad_idx.numel():
            self.flat_s.index_fill_(0, self._s_pad_idx, -1.0)

    def _upload_step(self, *, lr, mask_lr=100.0, betas=(0.9, 0.98), sparsity_weight=0.0, current_step=0, max_step=1, **_):
        """Per-step scalars of the captured graph -> pinned host words -> device (two tiny async copies)."""
        self.opt_step += 1
        golden = 0x9E3779B97F4A7C15
        for i in range(3):
            v = (self.seed + i + self.step_id * golden) & 0xFFFFFFFFFFFFFFFF
            self._seeds_host[i] = v - (1 << 64) if v >= (1 << 63) else v
        bc1 = 1.0 - betas[0] ** self.opt_step
        bc2 = math.sqrt(1.0 - betas[1] ** self.opt_step)
        anneal = (1.0 + math.cos(min(1.0, current_step / max_step) * math.pi)) / 2.0
        h = self._dyn_host
        h[0], h[1], h[2], h[3], h[4], h[5], h[6] = lr, bc1, bc2, mask_lr, bc1, bc2, sparsity_weight * (1.0 - anneal)
        self._seeds_dev.copy_(self._seeds_host, non_blocking=True)
        self._dyn_dev.copy_(self._dyn_host, non_blocking=True)

    def _graph_body(self, ws, part, opt):
        self._wm_step = {}  # every body (eager warm-up, capture) re-applies the masks
        if part == "all" and self.overlap_opt:
            # the decoder-side half of the optimizer (64 % of the parameters) runs on its own stream underneath the encoder's
            # backward: its gradients are final after phase 1 and nothing the remaining phases read is touched
            self.forward(ws)
            self.backward_phase(ws, 0)
            self.backward_phase(ws, 1)
            main, ost = torch.cuda.current_stream(self.dev), self._opt_stream()
            ost.wait_stream(main)
            with torch.cuda.stream(ost):
                self.optimizer_step(_dyn=True, part="hi", **opt)
            self.backward_phase(ws, 2)
            self.backward_phase(ws, 3)
            main.wait_stream(ost)
            self.optimizer_step(_dyn=True, part="lo", **opt)
            return
        if part in ("all", "fwdbwd"):
            self.forward(ws)
            self.loss_and_backward(ws)
        if part == "fwd_p0":
            self.forward(ws)
            self.backward_phase(ws, 0)
        if part in ("p1", "p2", "p3"):
            self.backward_phase(ws, int(part[1]))
        if part in ("all", "opt"):
            self.optimizer_step(_dyn=True, **opt)
        if part in ("opt_hi", "opt_lo"):
            self.optimizer_step(_dyn=True, part=part[4:], **opt)

    def _exchange(self, all_reduce, phase, handles):
        """Start the all-reduce of every gradient bucket that `phase` finished (async: the next phase overlaps it)."""
        for flat, a, b in self.grad_buckets(phase):
            h = all_reduce(flat[a:b])
            if h is not None:
                handles.append((phase, h))

    @staticmethod
    def _wait(handles, phases):
        for ph, h in handles:
            if ph in phases:
                h.wait()

    def _run_graphed(self, ws, part, opt):
        key = (part, tuple(sorted((k, float(v)) for k, v in opt.items() if k in ("clip", "eps", "mask_eps", "weight_decay",
                                                                                "grad_scale", "sparsity_target"))))
        graphs = ws.__dict__.setdefault("graphs", {})
        if key not in graphs:
            # the first step of a shape runs eagerly (it is a real step and also warms every kernel up), then the same
            # launch sequence is captured for all later steps
            self._graph_body(ws, part, opt)
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            before = K.lib.launch_count
            with torch.cuda.graph(g):
                self._graph_body(ws, part, opt)
            graphs[key] = g
            # kernels of ours inside the captured graph: every replay launches this many (bench.py gpu_launches)
            ws.__dict__.setdefault("graph_launches", {})[key] = K.lib.launch_count - before
            return
        graphs[key].replay()
        K.lib.launch_count += ws.graph_launches[key]

    def train_step(self, att_feats, boxes, seqs, masks, att_masks=None, *, seq_per_img, lr, all_reduce=None, global_tokens=None,
                   exchange=None, **opt):
        """One full SMP step.  Data parallel: either ``all_reduce`` (callable applied to the gradient buckets, NCCL sum, every
        rank then runs the whole optimizer) or ``exchange`` (distributed.ShardedExchange: reduce-scatter -> Adam on the owned
        shard -> all-gather per bucket, overlapped with the backward)."""
        # programmatic dependent launch, measured on this chain (ms/step): the GEMM's EARLY trigger (bit 2: dependents may start
        # once its loads are issued) costs time here, PDL with the implicit trigger at exit helps: mask 2 5.42, 3 5.30, 7 5.36
        prev_pdl = K.set_pdl(self.pdl_mask)
        try:
            return self._train_step(att_feats, boxes, seqs, masks, att_masks, seq_per_img=seq_per_img, lr=lr, all_reduce=all_reduce,
                                    global_tokens=global_tokens, exchange=exchange, **opt)
        finally:
            K.set_pdl(prev_pdl)

    def _train_step(self, att_feats, boxes, seqs, masks, att_masks=None, *, seq_per_img, lr, all_reduce=None, global_tokens=None,
                    exchange=None, **opt):
        B, N = att_feats.shape[:2]
        T = seqs.shape[1] - 1
        ws = self._get_ws(B, N, seq_per_img, T, att_masks is not None)
        self.step_id += 1
        self.load_batch(ws, att_feats, boxes, seqs, masks, att_masks, global_tokens)
        if exchange is not None and self.training:
            if self.use_graph:
                self._upload_step(lr=lr, **opt)
                self.opt_step -= 1  # (_sharded_step counts the step itself)
            self._sharded_step(ws, exchange, dict(opt, lr=lr) if self.use_graph else opt, lr)
            return ws.loss_sum * ws.inv_norm
        if self.use_graph and self.training:
            self._upload_step(lr=lr, **opt)
            opt = dict(opt, lr=lr)
            if all_reduce is None:
                self._run_graphed(ws, "all", opt)
            else:
                # one graph per backward phase; the buckets a phase finished are exchanged while the next one computes
                # and the decoder-side half of the optimizer (its buckets arrived long ago) runs underneath the last buckets' exchange
                handles = []
                for phase, part in enumerate(("fwd_p0", "p1", "p2", "p3")):
                    self._run_graphed(ws, part, opt)
                    self._exchange(all_reduce, phase, handles)
                self._wait(handles, (0, 1))
                self._run_graphed(ws, "opt_hi", opt)
                self._wait(handles, (2, 3))
                self._run_graphed(ws, "opt_lo", opt)
            return ws.loss_sum * ws.inv_norm
        self.forward(ws)
        if all_reduce is None:
            self.loss_and_backward(ws)
        else:
            handles = []
            for phase in range(self.N_PHASES):
                self.backward_phase(ws, phase)
                self._exchange(all_reduce, phase, handles)
            self._wait(handles, (0, 1))
            self.optimizer_step(lr=lr, part="hi", **opt)
            self._wait(handles, (2, 3))
            self.optimizer_step(lr=lr, part="lo", **opt)
            return ws.loss_sum * ws.inv_norm
        self.optimizer_step(lr=lr, **opt)
        return ws.loss_sum * ws.inv_norm

    def state_dict(self):
        sd = {k: v.clone() for k, v in self.p.items()}
        sd.update({k + "_pruning_mask": v.clone() for k, v in self.s.items()})
        return sd


class ModuleTrainer:
    """``train_step`` for the configurations the fused engine does not lay out - ACORT weight sharing (``share_att_*`` /
    ``share_layer_*``, models/relation_transformer.py:80-89,162-176): the kernel-backed module tree under autograd (every
    forward / backward is a kernel, autograd accumulates the applications of a shared layer and draws a fresh Bernoulli mask per
    application like the reference), LanguageModelCriterion (utils/losses.py:32-43) + sparsity loss (pruning/prune.py:228-269),
    then clip_gradient + Adam for the two parameter groups (scripts/train_n_prune_transformer.py:67-82, utils/optim.py:116-126)
    as one ``sc_adam_clip`` launch per tensor.  It updates the module's own parameters (nothing to sync back)."""

    def __init__(self, model):
        self.model = model
        self.opt_step = 0
        self._mv = {}

    def train_step(self, att_feats, boxes, seqs, masks, att_masks=None, *, seq_per_img, lr, all_reduce=None, global_tokens=None,
                   mask_lr=100.0, clip=0.1, betas=(0.9, 0.98), eps=1e-9, mask_eps=1e-2, weight_decay=0.0, grad_scale=1.0,
                   sparsity_target=None, sparsity_weight=0.0, current_step=0, max_step=1):
        m = self.model
        dev = m._device()
        m.train()
        params = [(n, p) for n, p in m.named_parameters() if p.requires_grad]
        for _, p in params:
            p.grad = None
        seqs_d, masks_d = seqs.to(dev), masks.to(dev).float()
        assert seqs_d.shape[0] == att_feats.shape[0] * seq_per_img, (seqs_d.shape, att_feats.shape, seq_per_img)
        out = m(att_feats=att_feats.to(dev), boxes=boxes.to(dev), seqs=seqs_d, att_masks=None if att_masks is None else att_masks.to(dev))
        T = out.shape[1]
        target, w = seqs_d[:, 1: T + 1], masks_d[:, 1: T + 1]
        denom = w.sum() if global_tokens is None else global_tokens
        loss = -(out.gather(2, target.unsqueeze(2)).squeeze(2) * w).sum() / denom
        total = loss
        if getattr(m, "MASKED", False) and m.mask_type == "supermask" and sparsity_target is not None and sparsity_weight:
            total = total + m.compute_sparsity_loss(sparsity_target, sparsity_weight, current_step, max_step)
        total.backward()
        if all_reduce is not None:  # data parallel: sum the gradients of every rank (the loss is normalised by the global token count)
            flat = torch._utils._flatten_dense_tensors([p.grad for _, p in params])
            h = all_reduce(flat)
            if h is not None:
                h.wait()
            for (_, p), g in zip(params, torch._utils._unflatten_dense_tensors(flat, [p.grad for _, p in params])):
                p.grad.copy_(g)
        self.opt_step += 1
        with torch.no_grad():
            for n, p in params:
                is_logit = n.endswith("_pruning_mask")
                if n not in self._mv:
                    self._mv[n] = (torch.zeros_like(p, dtype=torch.float32).view(-1), torch.zeros_like(p, dtype=torch.float32).view(-1))
                mom, var = self._mv[n]
                g = p.grad.contiguous().view(-1)
                K.adam_clip(p.data.view(-1), g, mom, var, lr=mask_lr if is_logit else lr, betas=betas, eps=mask_eps if is_logit else eps,
                            weight_decay=0.0 if is_logit else weight_decay, clip=clip, grad_scale=grad_scale, step=self.opt_step)
                torch.autograd.graph.increment_version(p)  # (written through raw pointers: cached engines must notice)
        return loss.detach()

    def state_dict(self):
        return self.model.state_dict()

