"""Synthetic workload of SURVEY.md section 8d: 36-region x 2048-d bottom-up features with boxes, random-init
weights with the reference's parameter names, exactly floor(s*numel) zeros per prunable tensor
("randomly pruned weights").  Generated with torch on the target device (plumbing, not part of the timed path)."""
import math

import torch


def synthetic_inputs(B, N, feat, seed=8888, device="cpu", pin=False):
    g = torch.Generator(device="cpu").manual_seed(seed)
    att = torch.relu(torch.randn(B, N, feat, generator=g)) * 2.0  # bottom-up features are post-ReLU
    xy = torch.rand(B, N, 2, generator=g) * 0.7
    wh = torch.rand(B, N, 2, generator=g) * 0.25 + 0.05
    boxes = torch.cat((xy, torch.clamp(xy + wh, max=1.0)), -1)
    if pin:
        return att.pin_memory(), boxes.pin_memory()
    return att.to(device), boxes.to(device)


def random_state_dict(cfg, seed=1234, sparsity=0.0, device="cpu"):
    """Dense-class (`relation_transformer`) state dict: Xavier-uniform inner-model weights, default nn.Linear init
    for att_embed (sparse_caption/models/relation_transformer.py:329-338)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    d, ff, V, F, h = cfg.d_model, cfg.dim_feedforward, cfg.vocab_size, cfg.att_feat_size, cfg.num_heads
    sd = {}

    def xavier(o, i):
        return (torch.rand(o, i, generator=g) * 2 - 1) * math.sqrt(6.0 / (i + o))

    def lin(prefix, o, i):
        sd[prefix + ".weight"] = xavier(o, i)
        sd[prefix + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) / math.sqrt(i)

    def norm(prefix):
        sd[prefix + ".a_2"] = torch.ones(d)
        sd[prefix + ".b_2"] = torch.zeros(d)

    sd["att_embed.0.weight"] = (torch.rand(d, F, generator=g) * 2 - 1) / math.sqrt(F)
    sd["att_embed.0.bias"] = (torch.rand(d, generator=g) * 2 - 1) / math.sqrt(F)
    ne = 3 if cfg.share_att_encoder else 4
    nd = 3 if cfg.share_att_decoder else 4
    for which, n_att in (("encoder", ne), ("decoder", nd)):
        uids = cfg.uids("enc" if which == "encoder" else "dec")
        first = {}
        for pos, u in enumerate(uids):
            p = f"model.{which}.layers.{pos}"
            if u in first:  # shared layer: same tensors under every position
                src = f"model.{which}.layers.{first[u]}."
                for k in [k for k in sd if k.startswith(src)]:
                    sd[p + "." + k[len(src):]] = sd[k]
                continue
            first[u] = pos
            for j in range(n_att):
                lin(f"{p}.self_attn.linears.{j}", d, d)
            if which == "encoder":
                for j in range(h):
                    lin(f"{p}.self_attn.WGs.{j}", 1, 4 if cfg.no_box_trigonometric_embedding else 64)
            else:
                for j in range(n_att):
                    lin(f"{p}.src_attn.linears.{j}", d, d)
            lin(f"{p}.feed_forward.w_1", ff, d)
            lin(f"{p}.feed_forward.w_2", d, ff)
            for j in range(2 if which == "encoder" else 3):
                norm(f"{p}.sublayer.{j}.norm")
        norm(f"model.{which}.norm")
    sd["model.tgt_embed.0.lut.weight"] = xavier(V, d)
    lin("model.generator.proj", V, d)
    if sparsity > 0:
        seen = set()
        for k, w in sd.items():
            if k.endswith(".weight") and w.dim() == 2 and id(w) not in seen:
                seen.add(id(w))
                nz = int(sparsity * w.numel())
                if nz > 0:
                    w.view(-1)[torch.randperm(w.numel(), generator=g)[:nz]] = 0.0
    return {k: v.to(device) for k, v in sd.items()}
