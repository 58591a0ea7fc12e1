"""The reference's Python surface, exercised the way the reference's own callers use it:
the training-loop recipe through autograd, the step-wise decoding entry points, and the attention modules as modules.
All arithmetic runs in the CUDA kernels; expectations come from the golden fixtures (outputs of the imported reference)
and from the CPU oracle."""
import math

import pytest
import torch

from oracle import ort_oracle as O
from tests import golden_io

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _model(name, z, precision="fp32", train=False):
    import sparse_caption_b200.relation_transformer as R
    m = R.get_model(name)(z["cfg_dict"])
    m.load_state_dict(z["w"], strict=True)
    m = m.to(DEV)
    m.train(train)
    m.precision = precision
    return m


def _zero_dropout(m):
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0


def test_reference_training_recipe_matches_reference_gradients():
    """scripts/train_n_prune_transformer.py:136-153, verbatim control flow: model(**data) -> LanguageModelCriterion ->
    + compute_sparsity_loss -> loss.backward() -> clip_gradient -> optimizer.step(), on the pruned class in TRAIN mode with
    the Bernoulli uniforms of the reference run injected.  Gradients must equal the reference's autograd (golden fixture)."""
    from sparse_caption_b200 import sampler
    z = golden_io.load("ort_prune_tiny")
    m = _model("relation_transformer_prune", z, "fp32", train=True)
    _zero_dropout(m)
    named = dict(m.named_parameters())
    sampler.inject_uniforms({named[k + "_pruning_mask"]: u.to(DEV) for k, u in z["u"].items()})
    try:
        data = dict(att_feats=z["att_feats"].to(DEV), boxes=z["boxes"].to(DEV), seqs=z["seqs"].to(DEV))
        weights = [p for n, p in m.named_parameters() if not n.endswith("_pruning_mask")]
        masks = [p for n, p in m.named_parameters() if n.endswith("_pruning_mask")]
        optimizer = torch.optim.Adam([{"params": weights, "lr": 1e-3, "betas": (0.9, 0.98), "eps": 1e-9},
                                      {"params": masks, "lr": 100.0, "betas": (0.9, 0.98), "eps": 1e-2, "weight_decay": 0}])
        optimizer.zero_grad()
        out = m(**data)                                                       # [B*S, T, V] log-probs, autograd graph attached
        assert out.requires_grad and tuple(out.shape) == tuple(z["tf_logprobs_train"].shape)
        assert rel_err(out, z["tf_logprobs_train"]) < 2e-4
        target, mask = z["seqs"][:, 1:].to(DEV), z["masks"][:, 1:].to(DEV)
        loss = -(out.gather(2, target.unsqueeze(2)).squeeze(2) * mask).sum() / mask.sum()   # utils/losses.py:32-43
        assert abs(float(loss) - float(z["tf_loss_train"])) < 2e-4 * abs(float(z["tf_loss_train"]))
        before = {n: p.detach().clone() for n, p in m.named_parameters()}
        loss.backward()
        for k, g in z["g"].items():
            assert named[k].grad is not None, k
            assert rel_err(named[k].grad, g) < 1e-3, k
        # every trainable parameter received a gradient (key-projection biases: analytically zero, still a tensor)
        assert all(p.grad is not None for p in m.parameters() if p.requires_grad)
        for group in optimizer.param_groups:                                  # utils/optim.py:187-191
            torch.nn.utils.clip_grad_value_(group["params"], 0.1)
        optimizer.step()
        moved = sum(int(not torch.equal(before[n], p.detach())) for n, p in m.named_parameters())
        assert moved > 0.9 * len(before)
        # the sparsity loss is differentiable through the module too (prune.py:228-269)
        optimizer.zero_grad()
        sl = m.compute_sparsity_loss(0.9, 5.0, 5, 10)
        sl.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in masks)
    finally:
        sampler.inject_uniforms(None)


@pytest.mark.parametrize("name,cls", [("ort_tiny", "relation_transformer"), ("ort_tiny_masks", "relation_transformer"),
                                      ("acort_tiny", "relation_transformer")])
def test_module_forward_and_gradients_match_reference(name, cls):
    """Dense class through the module tree (incl. ACORT share_att='kv' + share_layer and padded regions): teacher-forcing
    log-probs and the sampled gradients of the golden fixtures."""
    z = golden_io.load(name)
    m = _model(cls, z, "fp32", train=True)
    _zero_dropout(m)
    am = z.get("att_masks")
    out = m(att_feats=z["att_feats"].to(DEV), boxes=z["boxes"].to(DEV), seqs=z["seqs"].to(DEV),
            att_masks=None if am is None else am.to(DEV))
    assert rel_err(out, z["tf_logprobs"]) < 2e-4
    target, mask = z["seqs"][:, 1:].to(DEV), z["masks"][:, 1:].to(DEV)
    loss = -(out.gather(2, target.unsqueeze(2)).squeeze(2) * mask).sum() / mask.sum()
    assert abs(float(loss) - float(z["tf_loss"])) < 2e-4 * abs(float(z["tf_loss"]))
    loss.backward()
    named = dict(m.named_parameters())
    for k, g in z["g"].items():
        assert rel_err(named[k].grad, g) < 1e-3, k


def test_acort_pruned_class_trains_through_autograd():
    """ACORT (shared attention projections + shared layers) in the PRUNED class: the module path accumulates the gradients of
    every application of a shared layer; checked against autograd through the oracle with the same (binarized) masks."""
    z = golden_io.load("acort_tiny")
    import sparse_caption_b200.relation_transformer as R
    cfg = dict(z["cfg_dict"])
    m = R.get_model("relation_transformer_prune")(cfg)
    sd0 = m.state_dict()  # shared layers: the same tensor appears under every position it is applied at
    canon, by_ptr = {}, {}
    for k, v in sd0.items():
        canon[k] = by_ptr.setdefault(v.data_ptr(), k)
    g = torch.Generator().manual_seed(3)
    sd = {}
    for k, v in sd0.items():
        if canon[k] != k:
            sd[k] = sd[canon[k]]
        elif k.endswith("_pruning_mask"):
            sd[k] = torch.randn(v.shape, generator=g) * 2.0 + 0.5
        else:
            sd[k] = z["w"][k].clone()
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()  # eval: binarized masks (deterministic), gradients still flow straight-through
    m.precision = "fp32"
    out = m(att_feats=z["att_feats"].to(DEV), boxes=z["boxes"].to(DEV), seqs=z["seqs"].to(DEV))
    # oracle with autograd over the UNIQUE parameters
    W = {k: v.clone().requires_grad_(True) for k, v in sd.items() if canon[k] == k and v.is_floating_point()
         and not k.endswith("_pruning_mask") and not k.endswith(".pe")}
    S = {k[: -len("_pruning_mask")]: v.clone().requires_grad_(True) for k, v in sd.items() if canon[k] == k and k.endswith("_pruning_mask")}
    eff = {}
    for k in sd:
        if k.endswith("_pruning_mask") or k.endswith(".pe"):
            continue
        src = canon[k]
        w = W[src]
        if src in S:
            p = torch.sigmoid(S[src])
            w = (p + (O.binarize_logits(S[src].detach()) - p).detach()) * w
        eff[k] = w
    eff["model.tgt_embed.1.pe"] = sd["model.tgt_embed.1.pe"]
    lp = O.forward_tf(eff, z["cfg"], z["att_feats"], z["boxes"], z["seqs"], None)
    assert rel_err(out, lp) < 2e-4
    target, mask = z["seqs"][:, 1:], z["masks"][:, 1:]
    O.lm_criterion(lp, target, mask).backward()
    loss = -(out.gather(2, target.to(DEV).unsqueeze(2)).squeeze(2) * mask.to(DEV)).sum() / mask.to(DEV).sum()
    loss.backward()
    named = dict(m.named_parameters())  # (deduplicated: one entry per unique parameter, under its first position)
    checked = 0
    for n, p in named.items():
        ref = S[canon[n][: -len("_pruning_mask")]].grad if n.endswith("_pruning_mask") else W[canon[n]].grad
        if ref is None or float(ref.abs().max()) < 1e-8:
            continue
        assert rel_err(p.grad, ref) < 2e-3, n
        checked += 1
    assert checked > 40


def test_get_logprobs_state_and_batch_beam_search():
    """The step-wise entry points (relation_transformer.py:374-387, caption_model.py:30-226) used like
    CachedTransformerBase._generate_captions uses them (transformer.py:481-505), against the reference's beam search."""
    for name in ("ort_tiny", "acort_tiny"):
        z = golden_io.load(name)
        m = _model("relation_transformer", z, "fp32")
        c = m.cfg
        att, boxes = z["att_feats"].to(DEV), z["boxes"].to(DEV)
        B = att.shape[0]
        import sparse_caption_b200.relation_transformer as R
        with torch.no_grad(), R._Precision("fp32"):
            feats, bx, _, att_masks, _ = m._prepare_feature(att, None, boxes)
            memory = m.model.encode(feats, bx, att_masks)
        mem_ref, _ = O.encode(z["w"], z["cfg"], z["att_feats"], z["boxes"], None)
        assert rel_err(memory, mem_ref) < 2e-4
        for key, opt in (("beam3", {"beam_size": 3}), ("beam2", {"beam_size": 2}),
                         ("beam3c", {"beam_size": 3, "decoding_constraint": 1, "length_penalty": "wu_0.5"})):
            beam = opt["beam_size"]
            it = torch.full((B,), c.bos_token_id, dtype=torch.long, device=DEV)
            logprobs, state = m.get_logprobs_state(it, memory, att_masks, None)
            assert tuple(logprobs.shape) == (B, c.vocab_size)
            assert abs(float(logprobs.exp().sum(-1).mean()) - 1.0) < 1e-4
            mem_r, mask_r = torch.repeat_interleave(memory, beam, 0), torch.repeat_interleave(att_masks, beam, 0)
            done = m.batch_beam_search(state, logprobs, mem_r, mask_r, opt=opt)
            assert len(done) == B and all(len(d) == beam for d in done)
            for b in range(B):
                for v in range(beam):
                    want = z[key + "_seq"][b, v]
                    n = int((want != 0).sum())
                    assert done[b][v]["seq"].cpu().tolist() == want[:n].tolist(), (name, key, b, v)
                    torch.testing.assert_close(done[b][v]["logps"].cpu(), z[key + "_lp"][b, v, :n], rtol=1e-4, atol=2e-5)
                ps = [d["p"] for d in done[b]]
                assert ps == sorted(ps, reverse=True)


def test_box_attention_module_surface():
    """BoxMultiHeadedAttention.forward / static BoxRelationalEmbedding / static box_attention as modules."""
    import sparse_caption_b200.relation_transformer as R
    from sparse_caption_b200 import masked_layer as ML
    z = golden_io.load("ort_tiny")
    sd, ocfg = z["w"], z["cfg"]
    ML.set_precision("fp32")
    try:
        h, d = ocfg.num_heads, ocfg.d_model
        att = R.BoxMultiHeadedAttention(h, d, True, 0.0, None).to(DEV).eval()
        p = "model.encoder.layers.0.self_attn"
        att.load_state_dict({k[len(p) + 1:]: v for k, v in sd.items() if k.startswith(p + ".")})
        g = torch.Generator().manual_seed(1)
        x = torch.randn(3, 12, d, generator=g)
        boxes = z["boxes"]
        mask = torch.ones(3, 1, 12)
        mask[1, 0, 9:] = 0
        emb = R.BoxMultiHeadedAttention.BoxRelationalEmbedding(boxes.to(DEV))
        ref_emb = O.box_relational_embedding(boxes)
        assert tuple(emb.shape) == (3, 12, 12, 64)
        assert float((emb.cpu() - ref_emb).abs().max()) < 5e-5  # (sin / cos of angles up to 690 rad: fp32 argument rounding)
        y = att(x.to(DEV), x.to(DEV), x.to(DEV), boxes.to(DEV), mask.to(DEV))
        ref = O.box_mha(x, ref_emb, mask, sd, p, h)
        assert rel_err(y, ref) < 2e-4
        # static box_attention on [B, h, N, d_k] tensors with explicit relu'd geometry weights
        q, k, v = (torch.randn(3, h, 12, d // h, generator=g) for _ in range(3))
        wg = torch.relu(torch.randn(3, h, 12, 12, generator=g))
        out, probs = R.BoxMultiHeadedAttention.box_attention(q.to(DEV), k.to(DEV), v.to(DEV), wg.to(DEV), mask=mask.unsqueeze(1).to(DEV))
        sc = q @ k.transpose(-2, -1) / math.sqrt(d // h)
        sc = sc.masked_fill(mask.unsqueeze(1) == 0, -1e9)
        w = torch.softmax(torch.log(torch.clamp(wg, min=1e-6)) + sc, -1)
        assert rel_err(out, w @ v) < 1e-4 and rel_err(probs, w) < 1e-4
    finally:
        ML.set_precision("bf16")


def test_multi_headed_attention_cache_protocol():
    """MultiHeadedAttention.forward with .cache / .incremental_decoding / .reset_cache() (transformer.py:230-306): feeding a
    sequence token by token through the cached self-attention equals the masked full-sequence call."""
    import sparse_caption_b200.relation_transformer as R
    from sparse_caption_b200 import masked_layer as ML
    ML.set_precision("fp32")
    try:
        torch.manual_seed(0)
        h, d, T, B = 4, 64, 7, 3
        att = R.CachedMultiHeadedAttention(h, d, 0.0, True).to(DEV).eval()
        x = torch.randn(B, T, d, device=DEV)
        full_mask = torch.tril(torch.ones(1, T, T, device=DEV)).expand(B, T, T)
        full = att(x, x, x, full_mask)
        assert att.cache == [None, None]
        att.incremental_decoding = True
        att.reset_cache()
        steps = [att(x[:, t: t + 1], x[:, t: t + 1], x[:, t: t + 1], None) for t in range(T)]
        assert att.cache_size == 2 and tuple(att.cache[0].shape) == (B, h, T, d // h)
        assert rel_err(torch.cat(steps, 1), full) < 1e-4
        att.reset_cache()
        assert att.cache == [None, None]
        # cross-attention: projections of the memory are computed once and re-used; a smaller cached batch is repeated
        catt = R.CachedMultiHeadedAttention(h, d, 0.0, False).to(DEV).eval()
        catt.incremental_decoding = True
        mem = torch.randn(B, 5, d, device=DEV)
        y0 = catt(x[:, :1], mem, mem, None)
        k_cached = catt.cache[0]
        y1 = catt(x[:, :1].repeat_interleave(2, 0), mem.repeat_interleave(2, 0), mem.repeat_interleave(2, 0), None)
        assert catt.cache[0].shape[0] == 2 * B and torch.equal(catt.cache[0][::2], k_cached)
        assert rel_err(y1[::2], y0) < 1e-5
    finally:
        ML.set_precision("bf16")


def test_remove_bad_endings_and_suppress_unk():
    """caption_model.py:160-172: token 0 may not follow a 'bad ending'; UNK (last column) is lowered by 1000 when the model
    carries a vocabulary - both through OrtEngine.decode and against a restatement on the oracle's log-probs."""
    z = golden_io.load("ort_tiny")
    m = _model("relation_transformer", z, "fp32")
    att, boxes = z["att_feats"].to(DEV), z["boxes"].to(DEV)
    with pytest.raises(AttributeError):
        m(att_feats=att, boxes=boxes, opt={"beam_size": 3, "remove_bad_endings": 1}, mode="sample")  # no bad_endings_ix, like the reference
    base, _ = m(att_feats=att, boxes=boxes, opt={"beam_size": 3}, mode="sample")
    # make the most frequent non-final token of the plain search a "bad ending" and its successor-to-be token 0 attractive
    m.bad_endings_ix = [int(base[0, 0, 0])]
    V = m.cfg.vocab_size
    m.vocab = {str(V - 1): "UNK"}
    with torch.no_grad():
        m.model.generator.proj.bias[0] += 6.0       # token 0 would win everywhere ...
        m.model.generator.proj.bias[V - 1] += 5.0   # ... and UNK right behind it
    free, _ = m(att_feats=att, boxes=boxes, opt={"beam_size": 3}, mode="sample")
    assert int((free == 0).sum()) > 0 and int((free[:, 0] == V - 1).sum()) >= 0
    seq, lp = m(att_feats=att, boxes=boxes, opt={"beam_size": 3, "remove_bad_endings": 1, "suppress_UNK": 1}, mode="sample")
    bad = m.bad_endings_ix[0]
    for b in range(seq.shape[0]):
        for v in range(seq.shape[1]):
            s = seq[b, v].tolist()
            for t in range(1, len(s)):
                if s[t - 1] == bad and t < len(s):
                    assert not (s[t] == 0 and any(x != 0 for x in s[t:])), (b, v, s)  # 0 here could only be padding
            assert V - 1 not in s, (b, v, s)
    # the oracle with the same rule restated on its log-probs
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    memory, sm = O.encode(sd, z["cfg"], z["att_feats"], z["boxes"], None)
    ref = _oracle_beam_with_rules(sd, z["cfg"], memory, sm, 3, [bad], V - 1)
    assert torch.equal(seq.cpu(), ref)


def _oracle_beam_with_rules(sd, cfg, memory, src_mask, beam, bad_ix, unk_col):
    """beam_search of the oracle with remove_bad_endings / suppress_UNK applied to the log-probs (caption_model.py:160-172)."""
    B, L, V = memory.size(0), cfg.max_seq_length, cfg.vocab_size
    st = O.DecodeState(cfg)
    it = torch.full((B,), cfg.bos_token_id, dtype=torch.long)
    logprobs = O.decode_step(sd, cfg, st, it, memory, src_mask)
    mem_r, mask_r = memory.repeat_interleave(beam, 0), src_mask.repeat_interleave(beam, 0)
    beam_seq = torch.zeros(B, beam, 0, dtype=torch.long)
    beam_sum = torch.zeros(B, beam)
    done = [[] for _ in range(B)]
    for t in range(L):
        logprobs = logprobs.clone()
        if t > 0:
            prev = beam_seq[:, :, t - 1].reshape(-1)
            rows = torch.isin(prev, torch.tensor(bad_ix))
            logprobs[rows, 0] = float("-inf")
        logprobs[:, unk_col] -= 1000
        parent, word, new_sum = O.beam_select(logprobs, beam_sum, beam, first=(t == 0))
        nb = logprobs.size(0) // B
        if t > 0:
            beam_seq = beam_seq.gather(1, parent.unsqueeze(-1).expand_as(beam_seq))
        beam_seq = torch.cat([beam_seq, word.unsqueeze(-1)], -1)
        beam_sum = new_sum.clone()
        st.reorder((parent + torch.arange(B).unsqueeze(-1) * nb).reshape(-1))
        for b in range(B):
            is_end = beam_seq[b, :, t] == cfg.eos_token_id
            if t == L - 1:
                is_end = torch.ones_like(is_end)
            for v in range(beam):
                if is_end[v]:
                    done[b].append({"seq": beam_seq[b, v].clone(), "p": float(beam_sum[b, v])})
            beam_sum[b, is_end] -= 1000
        logprobs = torch.log_softmax(O.decode_step(sd, cfg, st, beam_seq[:, :, t].reshape(-1), mem_r, mask_r), -1)
    seq = torch.zeros(B, beam, L, dtype=torch.long)
    for b in range(B):
        for v, d in enumerate(sorted(done[b], key=lambda d: -d["p"])[:beam]):
            seq[b, v, : d["seq"].numel()] = d["seq"]
    return seq


def test_scst_style_rollout_in_train_mode_draws_bernoulli_masks():
    """utils/training.py:224-237: SCST samples with the model in train mode - the pruned class then decodes with a Bernoulli
    mask sample (one draw per call here), not the binarized masks."""
    from sparse_caption_b200 import sampler
    z = golden_io.load("ort_prune_tiny")
    m = _model("relation_transformer_prune", z, "fp32", train=True)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("_pruning_mask"):
                p.zero_()  # keep-probability 0.5 everywhere: the sample matters
    att, boxes = z["att_feats"].to(DEV), z["boxes"].to(DEV)
    sampler.set_mask_seed(5)
    a, _ = m(att_feats=att, boxes=boxes, opt={"beam_size": 2}, mode="sample")
    b_, _ = m(att_feats=att, boxes=boxes, opt={"beam_size": 2}, mode="sample")
    sampler.set_mask_seed(5)
    c, _ = m(att_feats=att, boxes=boxes, opt={"beam_size": 2}, mode="sample")
    assert torch.equal(a, c) and not torch.equal(a, b_)
    m.eval()
    e1, _ = m(att_feats=att, boxes=boxes, opt={"beam_size": 2}, mode="sample")
    e2, _ = m(att_feats=att, boxes=boxes, opt={"beam_size": 2}, mode="sample")
    assert torch.equal(e1, e2)


def test_acort_trainer_step_matches_reference_recipe():
    """``model.trainer()`` of an ACORT configuration (shared attention projections + shared layers, pruned class): one
    ``train_step`` = the reference recipe (loss.backward() -> clip_gradient(0.1) -> Adam with the two parameter groups of
    scripts/train_n_prune_transformer.py:67-82) applied to the gradients the kernel-backed module tree produced; shared layers
    stay shared; the loss of a repeated batch goes down."""
    from sparse_caption_b200.trainer import ModuleTrainer
    z = golden_io.load("acort_tiny")
    import sparse_caption_b200.relation_transformer as R
    m = R.get_model("relation_transformer_prune")(dict(z["cfg_dict"])).to(DEV)
    m.precision = "fp32"
    tr = m.trainer()
    assert isinstance(tr, ModuleTrainer) and m.trainer() is tr

    def shared_ok():
        for stack, share in ((m.model.encoder, m.cfg.share_layer_encoder), (m.model.decoder, m.cfg.share_layer_decoder)):
            for i, u in enumerate(share):
                for j, v in enumerate(share):
                    if (stack.layers[i] is stack.layers[j]) != (u == v):
                        return False
        return True
    assert m.cfg.share_layer_encoder and m.cfg.share_layer_decoder and shared_ok()
    named = [(n, p) for n, p in m.named_parameters() if p.requires_grad]
    before = {n: p.detach().clone() for n, p in named}
    S = z["seqs"].shape[0] // z["att_feats"].shape[0]
    kw = dict(seq_per_img=S, lr=1e-3, mask_lr=10.0, sparsity_target=0.9, sparsity_weight=1.0, current_step=500, max_step=1000)
    loss0 = float(tr.train_step(z["att_feats"], z["boxes"], z["seqs"], z["masks"], **kw))
    # the same update with torch: clip_gradient (clamp to +-0.1, utils/optim.py:187-191) then Adam, first step
    for n, p in named:
        g = p.grad.clamp(-0.1, 0.1)
        logit = n.endswith("_pruning_mask")
        lr, eps = (10.0, 1e-2) if logit else (1e-3, 1e-9)
        mhat, vhat = g, g * g                          # (1 - b) g / (1 - b^1) = g
        ref = before[n] - lr * mhat / (vhat.sqrt() + eps)
        assert rel_err(p.detach(), ref) < 1e-5, n
    assert shared_ok()
    losses = [float(tr.train_step(z["att_feats"], z["boxes"], z["seqs"], z["masks"], **kw)) for _ in range(12)]
    assert math.isfinite(loss0) and min(losses[-3:]) < loss0, (loss0, losses)
    m.sync_from_trainer()  # no-op for the module trainer

