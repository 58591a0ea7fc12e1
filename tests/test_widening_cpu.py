"""SURVEY.md section 8f rows, CPU tier: radix detokenisation against the IMPORTED reference tokenizer's outputs (golden
fixture), the .npy -> padded pinned batch path against the reference collate's semantics, checkpoint flavours."""
import json
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_radix_detok_matches_reference_tokenizer_golden():
    """detok.radix_to_word_ids == RadixTokenizer._decode_radix_ids (tokenizer.py:583-602) run from the imported reference on 3
    radix bases x 96 captions: with / without <eos>, ragged last groups, garbage after <eos>, pads."""
    from sparse_caption_b200.detok import radix_to_word_ids
    z = np.load(os.path.join(GOLDEN, "radix_detok.npz"))
    for name in ("b768", "b256", "b32"):
        base, tpw, eos, bos = (int(x) for x in z[f"{name}_meta"])
        seq = torch.from_numpy(z[f"{name}_seq"])
        want = torch.from_numpy(z[f"{name}_words"])
        words, n = radix_to_word_ids(seq, base, tpw, eos_id=eos)
        assert words.shape == want.shape
        assert torch.equal(n, (want >= 0).sum(-1))
        assert torch.equal(torch.where(want >= 0, words, torch.full_like(words, -1)), want)
        # the [B, beam, L] form the engine returns
        w3, n3 = radix_to_word_ids(seq.view(32, 3, -1), base, tpw, eos_id=eos)
        assert torch.equal(w3.view(96, -1), words) and torch.equal(n3.view(-1), n)


def test_detokenize_and_json_dump(tmp_path):
    from sparse_caption_b200 import evaluate as E
    z = np.load(os.path.join(GOLDEN, "radix_detok.npz"))
    base, tpw, eos, _ = (int(x) for x in z["b768_meta"])
    seq = torch.from_numpy(z["b768_seq"][:8])
    want = z["b768_words"][:8]
    vocab = {i: f"w{i}" for i in range(10000)}
    dec = lambda ids: " ".join(vocab[i] for i in ids)
    caps = E.detokenize(seq, decode_words=dec, radix=(base, tpw), vocab_len=10000)
    for c, w in zip(caps, want):
        assert c == " ".join(f"w{int(i) if int(i) < 10000 else 1}" for i in w[w >= 0])  # beyond the vocabulary -> <unk> (tokenizer.py:640)
    # word-level ids: cut at <eos>, drop pads
    s = torch.tensor([[5, 6, 7, 3, 9, 0], [8, 8, 0, 0, 0, 0]])
    assert E.detokenize(s, decode_words=dec) == ["w5 w6 w7", "w8 w8"]
    path = E.coco_caption_json_dump(zip([11, 12], ["a cat", "a dog"]), str(tmp_path / "val_beam_3" / "caption_00000001.json"))
    assert json.load(open(path)) == [{"image_id": 11, "caption": "a cat"}, {"image_id": 12, "caption": "a dog"}]
    with pytest.raises(AssertionError):
        E.coco_caption_json_dump([], str(tmp_path / "x.txt"))


def test_feature_batcher_matches_reference_collate_semantics(tmp_path):
    """.npy features / boxes of variable N -> zero-padded batch + masks, as ObjectRelationCollate builds them (collate.py:107-131,
    196-216: float32 cast, reshape(-1, F), pad_sequence(padding_value=0), att_masks = ones up to N_i)."""
    from sparse_caption_b200.ingest import FeatureBatcher
    rng = np.random.RandomState(0)
    att_dir, box_dir = tmp_path / "cocobu_att", tmp_path / "cocobu_box_relative"
    att_dir.mkdir(), box_dir.mkdir()
    ns = [22, 47, 36, 10, 47]
    for i, n in enumerate(ns):
        np.save(att_dir / f"{100 + i}.npy", (rng.rand(n, 64) * 30).astype(np.float32))
        np.save(box_dir / f"{100 + i}.npy", rng.rand(n, 4).astype(np.float64))       # the fixture's boxes are float64
    fb = FeatureBatcher(str(att_dir), str(box_dir), feat_dim=64, max_boxes=100, max_batch=8, pin=False)
    out = fb.from_ids([100 + i for i in range(5)])
    atts = [torch.from_numpy(np.load(att_dir / f"{100 + i}.npy").astype("float32")) for i in range(5)]
    boxes = [torch.from_numpy(np.load(box_dir / f"{100 + i}.npy").astype("float32")) for i in range(5)]
    want_att = torch.nn.utils.rnn.pad_sequence(atts, batch_first=True, padding_value=0.0)
    want_box = torch.nn.utils.rnn.pad_sequence(boxes, batch_first=True, padding_value=0.0)
    want_mask = torch.nn.utils.rnn.pad_sequence([torch.ones(n) for n in ns], batch_first=True, padding_value=0.0)
    assert out["att_feats"].is_contiguous() and torch.equal(out["att_feats"], want_att)
    assert torch.equal(out["boxes"], want_box) and torch.equal(out["att_masks"], want_mask)
    # a second batch re-uses the other staging buffer: the first result stays intact (double buffering)
    out2 = fb.from_ids([102, 102])
    assert out2["att_masks"] is None and tuple(out2["att_feats"].shape) == (2, 36, 64)   # fixed N: no mask needed
    assert torch.equal(out["att_feats"], want_att)
    # bf16 staging halves the H2D bytes
    fb16 = FeatureBatcher(str(att_dir), str(box_dir), feat_dim=64, max_boxes=47, max_batch=8, dtype=torch.bfloat16, pin=False)
    o16 = fb16.from_ids([100, 101])
    assert o16["att_feats"].dtype == torch.bfloat16 and torch.equal(o16["att_feats"].float(), want_att[:2].bfloat16().float())


def test_checkpoint_flavours_round_trip(tmp_path):
    """state_dict_sparse / _dense / _bin_mask files of a pruned model (prune.py:176-226) are recognised and reduced to the same
    dense-class weights; COO entries survive torch.save / torch.load."""
    import sparse_caption_b200.relation_transformer as R
    from sparse_caption_b200 import checkpoint as C, prune
    cfg = dict(d_model=32, dim_feedforward=64, num_layers=2, num_heads=4, max_seq_length=8, att_feat_size=48, vocab_size=37,
               prune_type="supermask", prune_supermask_init=5.0)
    torch.manual_seed(0)
    m = R.get_model("relation_transformer_prune")(cfg)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("_pruning_mask"):
                p.copy_(torch.randn_like(p) * 3)
    raw = {k: v.clone() for k, v in m.state_dict().items()}
    files = C.save_pruned_checkpoints(m, str(tmp_path))
    assert set(files) == {"sparse", "dense", "bin_mask"} and os.path.exists(tmp_path / "sparsities.csv")
    want = prune.fold_masks(raw, "supermask")
    kinds = {}
    for name, path in files.items():
        sd, kind = C.to_dense_class(C.load_checkpoint(path))
        kinds[name] = kind
        assert not any(k.endswith("_pruning_mask") for k in sd)
        for k, v in want.items():
            got = sd[k].to_dense() if sd[k].is_sparse else sd[k]
            assert torch.equal(got, v), (name, k)
    assert kinds == {"sparse": "sparse", "dense": "dense", "bin_mask": "bin_mask"}
    assert C.classify(raw) == "supermask"
    sd, _ = C.to_dense_class(raw)
    assert all(torch.equal(sd[k], want[k]) for k in want)
    # the sparse file is the small one (weights at ~50 % sparsity here; COO = values + 2 int64 indices)
    assert any(v.is_sparse for v in C.load_checkpoint(files["sparse"]).values())
