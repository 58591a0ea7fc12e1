"""Generate golden vectors by running the UNMODIFIED reference (imported read-only from
/root/reference) on seeded inputs and weights.  Run in the build container only:

    python tests/golden/make_golden.py

The outputs (``tests/golden/*.npz``) are committed; the GPU box never sees /root/reference.
Weights come from ``oracle.ort_oracle.random_state_dict`` and are stored in the fixture, so the
fixture is self-contained and independent of any RNG implementation.
"""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from sparse_caption.models import get_model  # noqa: E402
from sparse_caption.utils.config import Config  # noqa: E402
from sparse_caption.models.relation_transformer import BoxMultiHeadedAttention  # noqa: E402

from oracle import ort_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def ref_config(**kw):
    base = dict(share_att_encoder=None, share_att_decoder=None, share_layer_encoder=None, share_layer_decoder=None,
                d_model=64, dim_feedforward=128, num_layers=2, num_heads=4, drop_prob_src=0.5, max_seq_length=8,
                att_feat_size=96, vocab_size=50, eos_token_id=3, bos_token_id=2, unk_token_id=1, pad_token_id=0,
                no_box_trigonometric_embedding=False, prune_type="supermask", prune_mask_freeze_scope="",
                prune_supermask_init=5.0)
    base.update(kw)
    return base


def to_np(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}


def save(name, cfg, sd, extra):
    blob = {"cfg_keys": np.array(list(cfg.keys())), "cfg_vals": np.array([repr(v) for v in cfg.values()])}
    # the positional-encoding buffer is a pure function of d_model; it is rebuilt at load time
    blob.update({"w::" + k: v for k, v in to_np(sd).items() if not k.endswith(".pe")})
    blob.update(extra)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **blob)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def case_dense(name, B=3, N=12, S=2, lengths=None, eos_bias=0.0, **cfg_kw):
    cfg = ref_config(**cfg_kw)
    ocfg = O.Cfg(**cfg)
    sd = O.random_state_dict(ocfg, seed=1234)
    sd["model.generator.proj.bias"][cfg["eos_token_id"]] += eos_bias
    model = get_model("relation_transformer")(Config(**cfg))
    model.load_state_dict(sd, strict=True)
    model.eval()
    data = O.synthetic_inputs(B, N, cfg["att_feat_size"], seed=8888, seq_per_img=S,
                              max_len=cfg["max_seq_length"] + 2, vocab=cfg["vocab_size"])
    att_masks = None
    if lengths is not None:
        att_masks = torch.zeros(B, N)
        for b, n in enumerate(lengths):
            att_masks[b, :n] = 1
            data["att_feats"][b, n:] = 0
            data["boxes"][b, n:] = 0
    extra = {"att_feats": data["att_feats"].numpy(), "boxes": data["boxes"].numpy(), "seqs": data["seqs"].numpy(),
             "masks": data["masks"].numpy()}
    if att_masks is not None:
        extra["att_masks"] = att_masks.numpy()
    with torch.no_grad():
        out = model(att_feats=data["att_feats"], boxes=data["boxes"], seqs=data["seqs"], att_masks=att_masks)
        extra["tf_logprobs"] = out.numpy()
        for beam in (3, 2):
            seq, lp = model(att_feats=data["att_feats"], boxes=data["boxes"], att_masks=att_masks,
                            opt={"beam_size": beam}, mode="sample")
            extra[f"beam{beam}_seq"] = seq.numpy()
            extra[f"beam{beam}_lp"] = lp.numpy()
        seq, lp = model(att_feats=data["att_feats"], boxes=data["boxes"], att_masks=att_masks,
                        opt={"beam_size": 3, "decoding_constraint": 1, "length_penalty": "wu_0.5"}, mode="sample")
        extra["beam3c_seq"], extra["beam3c_lp"] = seq.numpy(), lp.numpy()
        seq, lp = model(att_feats=data["att_feats"], boxes=data["boxes"], att_masks=att_masks,
                        opt={"beam_size": 1}, mode="sample")
        extra["greedy_seq"], extra["greedy_lp"] = seq.numpy(), lp.numpy()
    # teacher-forcing loss + a few gradients (dense weights)
    model.zero_grad()
    out = model(att_feats=data["att_feats"], boxes=data["boxes"], seqs=data["seqs"], att_masks=att_masks)
    from sparse_caption.utils.losses import LanguageModelCriterion
    loss = LanguageModelCriterion()(out, data["seqs"][:, 1:], data["masks"][:, 1:])
    loss.backward()
    extra["tf_loss"] = np.array(loss.item(), dtype=np.float32)
    grads = dict(model.named_parameters())
    for k in ("att_embed.0.weight", "model.encoder.layers.0.self_attn.WGs.1.weight",
              "model.encoder.layers.1.self_attn.linears.0.weight", "model.decoder.layers.0.src_attn.linears.1.weight",
              "model.decoder.layers.1.feed_forward.w_2.bias", "model.generator.proj.weight",
              "model.tgt_embed.0.lut.weight", "model.encoder.norm.a_2"):
        if k in grads:  # shared layers appear once in named_parameters
            extra["g::" + k] = grads[k].grad.numpy()
    save(name, cfg, sd, extra)


def case_prune(name, B=2, N=10, S=2):
    """relation_transformer_prune: eval-mode TF forward (binarized masks) and train-mode forward/backward
    with torch.bernoulli replaced by (u < p) on recorded uniforms, dropout off."""
    cfg = ref_config()
    ocfg = O.Cfg(**cfg)
    sd = O.random_state_dict(ocfg, seed=4321)
    model = get_model("relation_transformer_prune")(Config(**cfg))
    g = torch.Generator().manual_seed(99)
    full = dict(sd)
    for k in O.prunable_keys(sd):
        # logits straddling 0 incl. exact zeros and tiny positives (binarization edge cases)
        s = torch.randn(sd[k].shape, generator=g) * 2.0 + 0.5
        flat = s.view(-1)
        flat[0] = 0.0
        if flat.numel() > 3:
            flat[1] = 5e-8
            flat[2] = 1e-7
            flat[3] = -0.0
        full[k + "_pruning_mask"] = s
    model.load_state_dict(full, strict=True)
    data = O.synthetic_inputs(B, N, cfg["att_feat_size"], seed=777, seq_per_img=S,
                              max_len=cfg["max_seq_length"] + 2, vocab=cfg["vocab_size"])
    extra = {"att_feats": data["att_feats"].numpy(), "boxes": data["boxes"].numpy(), "seqs": data["seqs"].numpy(),
             "masks": data["masks"].numpy()}
    model.eval()
    with torch.no_grad():
        out = model(att_feats=data["att_feats"], boxes=data["boxes"], seqs=data["seqs"])
        extra["tf_logprobs_eval"] = out.numpy()
    # sparsity loss (prune.py:228-269)
    sl = model.compute_sparsity_loss(0.8, 7.5, 30, 100)
    extra["sparsity_loss"] = np.array(float(sl), dtype=np.float32)
    # train mode with injected uniforms
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    uniforms = {}
    name_of = {id(p): n for n, p in model.named_parameters()}
    ug = torch.Generator().manual_seed(5)
    orig = torch.bernoulli
    state = {"calls": []}

    def fake_bernoulli(p, *a, **k):
        u = torch.rand(p.shape, generator=ug)
        state["calls"].append(u)
        return (u < p).to(p.dtype)

    torch.bernoulli = fake_bernoulli
    try:
        model.zero_grad()
        out = model(att_feats=data["att_feats"], boxes=data["boxes"], seqs=data["seqs"])
        from sparse_caption.utils.losses import LanguageModelCriterion
        loss = LanguageModelCriterion()(out, data["seqs"][:, 1:], data["masks"][:, 1:])
        loss.backward()
    finally:
        torch.bernoulli = orig
    extra["tf_logprobs_train"] = out.detach().numpy()
    extra["tf_loss_train"] = np.array(loss.item(), dtype=np.float32)
    # call order of bernoulli == order of MaskedLinear/Embedding forward calls; record by shape-walk
    order = []
    order.append("att_embed.0.weight")
    for i in range(cfg["num_layers"]):
        p = f"model.encoder.layers.{i}"
        order += [f"{p}.self_attn.linears.{j}.weight" for j in range(3)]
        order += [f"{p}.self_attn.WGs.{j}.weight" for j in range(cfg["num_heads"])]
        order += [f"{p}.self_attn.linears.3.weight", f"{p}.feed_forward.w_1.weight", f"{p}.feed_forward.w_2.weight"]
    order.append("model.tgt_embed.0.lut.weight")
    for i in range(cfg["num_layers"]):
        p = f"model.decoder.layers.{i}"
        order += [f"{p}.self_attn.linears.{j}.weight" for j in range(4)]
        order += [f"{p}.src_attn.linears.{j}.weight" for j in range(4)]
        order += [f"{p}.feed_forward.w_1.weight", f"{p}.feed_forward.w_2.weight"]
    order.append("model.generator.proj.weight")
    assert len(order) == len(state["calls"]), (len(order), len(state["calls"]))
    for k, u in zip(order, state["calls"]):
        assert tuple(u.shape) == tuple(full[k].shape), (k, u.shape)
        extra["u::" + k] = u.numpy()
    params = dict(model.named_parameters())
    for k in ("att_embed.0.weight", "model.encoder.layers.0.self_attn.WGs.1.weight",
              "model.encoder.layers.1.self_attn.linears.0.weight", "model.decoder.layers.0.src_attn.linears.1.weight",
              "model.generator.proj.weight", "model.tgt_embed.0.lut.weight"):
        extra["g::" + k] = params[k].grad.numpy()
        extra["g::" + k + "_pruning_mask"] = params[k + "_pruning_mask"].grad.numpy()
    save(name, cfg, full, extra)


def case_box():
    g = torch.Generator().manual_seed(3)
    xy = torch.rand(2, 9, 2, generator=g) * 0.7
    wh = torch.rand(2, 9, 2, generator=g) * 0.25 + 0.05
    boxes = torch.cat((xy, torch.clamp(xy + wh, max=1.0)), -1)
    boxes[1, 3] = boxes[1, 2]  # identical boxes -> clamp(1e-3) path
    boxes[0, 8] = 0.0  # padded (all-zero) box
    emb = BoxMultiHeadedAttention.BoxRelationalEmbedding(boxes)
    q, k, v = (torch.randn(2, 4, 9, 16, generator=g) for _ in range(3))
    gw = torch.relu(torch.randn(2, 4, 9, 9, generator=g))
    mask = torch.ones(2, 1, 1, 9)
    mask[0, :, :, 8] = 0
    out, w = BoxMultiHeadedAttention.box_attention(q, k, v, gw, mask=mask, dropout=None)
    np.savez_compressed(os.path.join(OUT, "box_geometry.npz"), boxes=boxes.numpy(), emb=emb.numpy(), q=q.numpy(),
                        k=k.numpy(), v=v.numpy(), g=gw.numpy(), mask=mask.numpy(), out=out.numpy(), w=w.numpy())
    print("box_geometry ok")


def case_binarize():
    """Bit-exactness pin for rint(sigmoid(S)) around the 0.5 threshold (sampler.py:57-66)."""
    from sparse_caption.pruning.sampler import rounding_sigmoid
    s = torch.cat([torch.tensor([0.0, -0.0, 1e-9, 5e-8, 8.9e-8, 8.94e-8, 9e-8, 1e-7, 1.2e-7, 2e-7, -1e-7, -5e-8, 1e-3,
                                 -1e-3, 5.0, -5.0, 88.0, -88.0, 1e4, -1e4, float("inf"), float("-inf")]),
                   torch.linspace(-3e-7, 3e-7, 601)])
    m = rounding_sigmoid(s)
    np.savez_compressed(os.path.join(OUT, "binarize.npz"), s=s.numpy(), m=m.numpy())
    print("binarize ok; threshold:", float(s[m > 0].abs().min()))


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(4)
    case_box()
    case_binarize()
    case_dense("ort_tiny", eos_bias=2.5)
    case_dense("ort_tiny_masks", B=3, N=12, lengths=[12, 9, 7], eos_bias=2.5)
    case_dense("acort_tiny", B=2, N=10, eos_bias=2.0, share_att_encoder="kv", share_att_decoder="kv",
               share_layer_encoder=(0, 0, 1, 1), share_layer_decoder=(0, 0, 1, 1), num_layers=4, vocab_size=35,
               max_seq_length=10)
    case_prune("ort_prune_tiny")
