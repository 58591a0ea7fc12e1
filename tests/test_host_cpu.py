"""CPU tier: host-side logic of the drop-in classes (no kernels run): parameter naming vs the reference, PruningMixin
bookkeeping, mask schedulers, state-dict export, CSR packing, loud failure without CUDA."""
import pytest
import torch

from oracle import ort_oracle as O
from tests import golden_io


def _cfg(**kw):
    c = dict(d_model=64, dim_feedforward=128, num_layers=2, num_heads=4, drop_prob_src=0.5, max_seq_length=8, att_feat_size=96,
             vocab_size=50, eos_token_id=3, bos_token_id=2, unk_token_id=1, pad_token_id=0, prune_type="supermask",
             prune_mask_freeze_scope="", prune_supermask_init=5.0)
    c.update(kw)
    return c


def test_reference_checkpoints_load_strict():
    """Parameter names / shapes equal the reference's (fixtures store the reference's own state-dict keys)."""
    import sparse_caption_b200.relation_transformer as R
    for name, model in (("ort_tiny", "relation_transformer"), ("ort_prune_tiny", "relation_transformer_prune")):
        z = golden_io.load(name)
        m = R.get_model(model)(z["cfg_dict"])
        m.load_state_dict(z["w"], strict=True)
    z = golden_io.load("acort_tiny")
    m = R.get_model("relation_transformer")(z["cfg_dict"])
    m.load_state_dict(z["w"], strict=True)
    assert m.model.encoder.layers[0] is m.model.encoder.layers[1]  # share_layer (0,0,1,1)
    assert len(m.model.decoder.layers[0].self_attn.linears) == 3    # share_att 'kv'


def test_registry_and_errors():
    import sparse_caption_b200.relation_transformer as R
    assert set(R.MODEL_REGISTRY) >= {"relation_transformer", "relation_transformer_prune"}
    with pytest.raises(ValueError):
        R.get_model("nope")
    m = R.get_model("relation_transformer")(_cfg())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(att_feats=torch.zeros(1, 3, 96), boxes=torch.zeros(1, 3, 4), mode="sample", opt={})


def test_pruning_mixin_bookkeeping():
    import sparse_caption_b200.relation_transformer as R
    z = golden_io.load("ort_prune_tiny")
    m = R.get_model("relation_transformer_prune")(z["cfg_dict"])
    m.load_state_dict(z["w"], strict=True)
    masks = m.all_pruning_masks()
    weights = m.all_pruned_weights()
    assert len(masks) == len(weights) == len([k for k in z["w"] if k.endswith("_pruning_mask")])
    assert m.total_mask_params == sum(p.numel() for _, p in masks)
    sp, nnz, per, names = m.all_mask_sparsities
    logits = [p for _, p in masks]
    ref_nnz = sum(O.binarize_logits(p).sum() for p in logits)
    assert float(nnz) == float(ref_nnz)
    assert abs(float(sp) - (1 - float(ref_nnz) / m.total_mask_params)) < 1e-6
    # sparsity loss value + straight-through gradient against the oracle
    loss = m.compute_sparsity_loss(0.8, 7.5, 30, 100)
    ref, _ = O.sparsity_loss(logits, 0.8, 7.5, 30, 100)
    torch.testing.assert_close(loss.detach().float(), ref.float(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(loss.detach().float(), z["sparsity_loss"], rtol=1e-5, atol=1e-7)
    loss.backward()
    p0 = logits[0]
    sig = torch.sigmoid(p0.detach())
    import math
    anneal = (1 + math.cos(0.3 * math.pi)) / 2
    sign = 1.0 if 0.8 - float(sp) >= 0 else -1.0
    expect = sign / m.total_mask_params * 7.5 * (1 - anneal) * sig * (1 - sig)
    torch.testing.assert_close(p0.grad, expect, rtol=1e-4, atol=1e-12)
    # dense / sparse export round trip (prune.py:176-226, model_utils.py:110-118)
    dense = m.state_dict_dense(discard_pruning_mask=True, prune_weights=True)
    assert not any(k.endswith("_pruning_mask") for k in dense)
    eff = O.effective_state_dict(z["w"], "supermask")
    for k in dense:
        if k in eff:
            torch.testing.assert_close(dense[k], eff[k])
    sparse = m.state_dict_sparse()
    from sparse_caption_b200.prune import densify_state_dict
    back = densify_state_dict(sparse)
    for k in dense:
        torch.testing.assert_close(back[k], dense[k])
    dm = R.get_model("relation_transformer")(z["cfg_dict"])
    dm.load_state_dict(back, strict=True)  # eval_model.py:64-77 flow


@pytest.mark.parametrize("mask_type", ["mag_blind", "mag_uniform", "mag_dist", "snip"])
def test_one_shot_mask_updates(mask_type):
    """update_masks_once hits the target sparsity (reference tests/test_prune.py:111-117 tolerance 0.05)."""
    import sparse_caption_b200.relation_transformer as R
    torch.manual_seed(8888)
    m = R.get_model("relation_transformer_prune")(_cfg(prune_type=mask_type))
    assert float(m.all_mask_sparsities[0]) == 0.0
    if mask_type == "snip":
        for p in m.all_pruning_masks(named=False):
            p.grad = torch.rand_like(p)
    m.update_masks_once(0.7)
    assert abs(float(m.all_mask_sparsities[0]) - 0.7) < 0.05
    if mask_type == "mag_uniform":
        assert all(abs(float(s) - 0.7) < 0.05 for s in m.all_mask_sparsities[2])


def test_gradual_schedule():
    import sparse_caption_b200.relation_transformer as R
    m = R.get_model("relation_transformer_prune")(_cfg(prune_type="mag_grad_uniform"))
    seen = []
    for step in range(0, 3001, 500):
        m.update_masks_gradual(0.9, step, start_step=1000, prune_steps=2, prune_frequency=1000)
        seen.append(round(float(m.all_mask_sparsities[0]), 2))
    assert seen[0] == 0.0 and seen[2] == 0.0 and abs(seen[4] - 0.79) < 0.03 and abs(seen[6] - 0.9) < 0.03


def test_csr_pack_matches_reference_sparse_export():
    from sparse_caption_b200.kernels import CsrWeight
    g = torch.Generator().manual_seed(0)
    w = torch.randn(37, 64, generator=g) * (torch.rand(37, 64, generator=g) > 0.9)
    w[5] = 0
    csr = CsrWeight(w, torch.float32)
    coo = w.to_sparse().coalesce()  # what state_dict_sparse stores
    assert csr.nnz == coo.values().numel()
    assert torch.equal(csr.val, coo.values())
    assert torch.equal(csr.col.to(torch.int64) & 0xFFFF, coo.indices()[1])
    rows = torch.repeat_interleave(torch.arange(37), (csr.row_ptr[1:] - csr.row_ptr[:-1]).long())
    assert torch.equal(rows, coo.indices()[0])
    assert CsrWeight(coo, torch.float32).nnz == csr.nnz  # accepts the reference's COO tensors directly


def test_att_parts_and_cfg():
    from sparse_caption_b200.engine import ModelCfg, _att_parts, _parse_penalty
    assert _att_parts(None) == (0, 1, 2, 3) and _att_parts("kv") == (0, 1, 1, 2) and _att_parts("qk") == (0, 0, 1, 2)
    c = ModelCfg(_cfg(share_layer_decoder=(0, 0, 1, 1), num_layers=4))
    assert c.uids("dec") == [0, 0, 1, 1] and c.uids("enc") == [0, 1, 2, 3]
    assert _parse_penalty("wu_0.5") == (1, 0.5) and _parse_penalty("") == (0, 0.0) and _parse_penalty("avg_0") == (2, 0.0)
    with pytest.raises(ValueError):
        ModelCfg(dict(d_model=64))


def test_masked_layer_attributes_and_cpu_refusal():
    from sparse_caption_b200.masked_layer import MaskedEmbedding, MaskedLinear
    lin = MaskedLinear(16, 8, "supermask", 5.0)
    assert lin.weight_pruning_mask.shape == lin.weight.shape and lin.weight_pruning_mask.requires_grad
    assert float(lin.weight_pruning_mask.mean()) == 5.0 and lin.mask_trainable
    assert set(dict(lin.named_parameters())) == {"weight", "bias", "weight_pruning_mask"}
    frozen = MaskedLinear(16, 8, "mask_freeze", None)
    assert not frozen.weight_pruning_mask.requires_grad and float(frozen.weight_pruning_mask.mean()) == 1.0
    emb = MaskedEmbedding(10, 8, "snip", None)
    assert emb.mask_trainable and emb.weight_pruning_mask.requires_grad
    with pytest.raises(AssertionError):
        MaskedLinear(4, 4, "bogus", 1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lin(torch.zeros(2, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        emb(torch.zeros(2, dtype=torch.long))


def test_gradient_buckets_partition_the_flat_buffers():
    """Data-parallel exchange: the buckets of the three backward phases are disjoint and cover every weight and
    mask-logit gradient exactly once (OrtTrainer.grad_buckets; layout logic only, no kernel runs on the CPU)."""
    import torch
    from oracle import ort_oracle as O
    from sparse_caption_b200.engine import ModelCfg
    from sparse_caption_b200.trainer import OrtTrainer
    for L in (1, 2, 3, 6):
        cfg = dict(d_model=32, dim_feedforward=64, num_layers=L, num_heads=4, max_seq_length=8, att_feat_size=48, vocab_size=37)
        sd = O.random_state_dict(O.Cfg(**cfg), seed=3, sparsity=0.0)
        for fused in (True, False):
            # fused_st: only dWm (flat_gw) is exchanged - sc_adam_clip_st forms dW and dS; otherwise both flat buffers are
            tr = OrtTrainer(sd, ModelCfg(cfg), mask_type="supermask", precision="fp32", device="cpu", fused_st=fused)
            for flat in (tr.flat_gw, tr.flat_gs):
                cover = torch.zeros(flat.numel(), dtype=torch.int32)
                for phase in range(tr.N_PHASES):
                    for f, a, b in tr.grad_buckets(phase):
                        if f is flat:
                            assert 0 <= a < b <= flat.numel()
                            cover[a:b] += 1
                want = 0 if (fused and flat is tr.flat_gs) else 1
                assert int(cover.min()) == want and int(cover.max()) == want, (L, fused, cover.unique())
        # phase 0 must contain the generator, the last phase the first (att_embed) weight
        assert any(f is tr.flat_gw and b == tr.flat_gw.numel() for f, a, b in tr.grad_buckets(0))
        assert any(f is tr.flat_gw and a == 0 for f, a, b in tr.grad_buckets(tr.N_PHASES - 1))


def test_radix_detokenisation_matches_reference_arithmetic():
    """Vectorised radix -> word ids (detok.radix_to_word_ids) against a restatement of RadixTokenizer._decode_radix_ids
    (tokenizer.py:583-602, 684-712) on random captions: with / without <eos>, ragged last groups, bos/pad inside."""
    import itertools
    import torch
    from sparse_caption_b200.detok import radix_to_word_ids

    def ref_decode(ids, base, tpw, eos):
        ids = list(ids)
        if eos in ids:
            ids = ids[: ids.index(eos)]
        groups = list(itertools.zip_longest(fillvalue=1, *([iter(ids)] * tpw)))
        return [sum(max(d - 1, 0) * base ** i for i, d in enumerate(reversed(g))) + 4 for g in groups]

    g = torch.Generator().manual_seed(0)
    for base, tpw, L in ((768, 2, 26), (256, 2, 17), (16, 3, 20), (768, 1, 9)):
        eos = base + 2
        seq = torch.randint(0, base + 1, (7, 3, L), generator=g)
        for b in range(7):
            for k in range(3):
                cut = int(torch.randint(0, L + 3, (1,), generator=g))
                if cut < L:
                    seq[b, k, cut] = eos
                    seq[b, k, cut + 1:] = 0
        seq[0, 0, 0] = base + 1  # a stray <bos> digit is decoded like any other id (as the reference does)
        words, n = radix_to_word_ids(seq, base, tpw)
        for b in range(7):
            for k in range(3):
                want = ref_decode(seq[b, k].tolist(), base, tpw, eos)
                assert int(n[b, k]) == len(want)
                assert words[b, k, : len(want)].tolist() == want
                assert bool((words[b, k, len(want):] == 0).all())


def test_sell_weight_packing_roundtrip():
    """Sliced-ELL packing (kernels.SellWeight, host-side tensor code): every non-zero lands at slab_ptr[s] + 32 i + lane with
    its column and value, slab widths are multiples of 4, padding entries are (column 0, value 0)."""
    import torch
    from sparse_caption_b200.kernels import SellWeight
    g = torch.Generator().manual_seed(0)
    for dt in (torch.bfloat16, torch.float32):
        N, Kd = 100, 72
        w = torch.randn(N, Kd, generator=g) * (torch.rand(N, Kd, generator=g) > 0.9)
        w[5] = 0
        if dt == torch.bfloat16:
            w = w.bfloat16().float()
        sw = SellWeight(w, dt)
        assert sw.nnz == int((w != 0).sum()) and sw.shape == (N, Kd)
        slabs = (N + 31) // 32
        assert sw.slab_ptr.numel() == slabs + 1 and int(sw.slab_ptr[0]) == 0
        rec = torch.zeros(slabs * 32, Kd)
        for s_ in range(slabs):
            b, e = int(sw.slab_ptr[s_]), int(sw.slab_ptr[s_ + 1])
            assert (e - b) % (32 * 4) == 0
            ent = sw.entries[b:e]
            if dt == torch.bfloat16:
                u = ent.long() & 0xFFFFFFFF
                cols = u >> 16
                vals = (u & 0xFFFF).to(torch.int32).to(torch.int16).view(torch.bfloat16).float()
            else:
                cols = ent[:, 0].long()
                vals = ent[:, 1].contiguous().view(torch.float32)
            lane = torch.arange(e - b) % 32
            rec.index_put_((s_ * 32 + lane, cols), vals, accumulate=True)
        assert torch.equal(rec[:N], w)


def test_decode_grid_hint():
    """Grid sizing of the decode GEMMs in the throughput regime (engine.decode_grid_hint): tiles per persistent CTA so that the
    wide GEMMs land on ~48 CTAs and the N = d_model ones on ~60; small batches keep the kernel's own heuristic."""
    from sparse_caption_b200.engine import decode_grid_hint
    dc = (48, 60)
    # 5 coalesced batches x 512 images x beam 3 = 7680 rows: qkv 60 x 6 = 360 tiles, ff1 480, N = 512 GEMMs 120
    assert decode_grid_hint(7680, 1536, 512, dc) == 8 * 10000000 + 3256
    assert decode_grid_hint(7680, 2048, 512, dc) == 10 * 10000000 + 3256
    assert decode_grid_hint(7680, 512, 512, dc) == 2 * 10000000 + 3256
    for rows, n in ((7680, 1536), (7680, 2048), (7680, 512), (12800, 1024), (30720, 2048)):
        hint = decode_grid_hint(rows, n, 512, dc)
        tiles = -(-rows // 128) * -(-n // 256)
        ctas = -(-tiles // (hint // 10000000))
        assert ctas <= (60 if n <= 512 else 48) and hint % 10000000 == 3256
    # one batch of 16 images: 1 M block, nothing to cap
    assert decode_grid_hint(48, 1536, 512, dc) == 0 and decode_grid_hint(1536, 512, 512, dc) == 0
    assert decode_grid_hint(10 ** 7, 2048, 512, dc) // 10000000 == 200   # (the hint's tiles-per-CTA digit is capped)


def test_acort_trainer_refuses_cpu():
    """``model.trainer()`` of an ACORT configuration hands out the module trainer; like every other entry of the path it fails
    loudly on a CPU model instead of falling back."""
    import sparse_caption_b200.relation_transformer as R
    from sparse_caption_b200.trainer import ModuleTrainer
    z = golden_io.load("acort_tiny")
    m = R.get_model("relation_transformer_prune")(dict(z["cfg_dict"]))
    tr = m.trainer()
    assert isinstance(tr, ModuleTrainer) and m.trainer() is tr
    S = z["seqs"].shape[0] // z["att_feats"].shape[0]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tr.train_step(z["att_feats"], z["boxes"], z["seqs"], z["masks"], seq_per_img=S, lr=1e-3)

