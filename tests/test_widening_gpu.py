"""SURVEY.md section 8f rows on the GPU: reference-format checkpoint files straight into the engine (f2), .npy bottom-up
features of variable N through the pinned batcher into the engine (f3), and the eval caller with radix detokenisation (f1)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ort_oracle as O
from tests import golden_io

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _pruned_model(z):
    import sparse_caption_b200.relation_transformer as R
    m = R.get_model("relation_transformer_prune")(z["cfg_dict"])
    m.load_state_dict(z["w"], strict=True)
    return m.to(DEV).eval()


@pytest.mark.parametrize("backend", ["dense", "sell", "csr"])
def test_checkpoint_files_into_engine(tmp_path, backend):
    """model_best_pruned_sparse.pth (COO) / _pruned.pth / _bin_mask.pth written from a pruned model, loaded back through
    checkpoint.engine_from_checkpoint: identical captions from all three, equal to the oracle on the folded weights - the
    reference's own final-evaluation flow (train_n_prune_transformer.py:293-301, eval_model.py:64-88)."""
    from sparse_caption_b200 import checkpoint as C
    z = golden_io.load("ort_prune_tiny")
    m = _pruned_model(z)
    eff = O.effective_state_dict(z["w"], "supermask")
    files = C.save_pruned_checkpoints(m, str(tmp_path))
    rseq, rlp = O.sample(eff, z["cfg"], z["att_feats"], z["boxes"], None, {"beam_size": 3})
    for name, path in files.items():
        eng = C.engine_from_checkpoint(path, z["cfg_dict"], device=DEV, precision="fp32", sparse_backend=backend)
        seq, lp = eng.sample(z["att_feats"], z["boxes"], None, {"beam_size": 3})
        assert torch.equal(seq.cpu(), rseq), (name, backend)
        torch.testing.assert_close(lp.cpu(), rlp, rtol=1e-4, atol=2e-5)
    # and the file loads into the reference-shaped dense class with strict=True (key contract)
    import sparse_caption_b200.relation_transformer as R
    dense = R.get_model("relation_transformer")(z["cfg_dict"])
    from sparse_caption_b200.prune import densify_state_dict
    sd = densify_state_dict(C.load_checkpoint(files["sparse"]))
    missing = dense.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys


def test_npy_batches_through_the_engine(tmp_path):
    """Variable-N .npy features + boxes -> FeatureBatcher (pinned, zero padded, masks) -> OrtEngine.submit: token-exact against
    the oracle on the same padded batch; bf16 staging feeds the bf16 engine."""
    from sparse_caption_b200.engine import ModelCfg, OrtEngine
    from sparse_caption_b200.ingest import FeatureBatcher
    z = golden_io.load("ort_tiny")
    F = z["cfg_dict"]["att_feat_size"]
    rng = np.random.RandomState(1)
    att_dir, box_dir = tmp_path / "att", tmp_path / "box"
    att_dir.mkdir(), box_dir.mkdir()
    ns = [12, 7, 10, 12, 5, 9]
    for i, n in enumerate(ns):
        np.save(att_dir / f"{i}.npy", np.maximum(rng.randn(n, F), 0).astype(np.float32) * 2)
        xy = rng.rand(n, 2) * 0.7
        np.save(box_dir / f"{i}.npy", np.concatenate([xy, np.minimum(xy + rng.rand(n, 2) * 0.25 + 0.05, 1.0)], 1))
    fb = FeatureBatcher(str(att_dir), str(box_dir), feat_dim=F, max_boxes=16, max_batch=8)
    batch = fb.from_ids(range(len(ns)))
    assert batch["att_feats"].is_pinned() and batch["att_masks"] is not None
    eng = OrtEngine(z["w"], ModelCfg(z["cfg_dict"]), precision="fp32", device=DEV)
    L = z["cfg_dict"]["max_seq_length"]
    out = (torch.zeros(len(ns), 3, L, dtype=torch.int32).pin_memory(), torch.zeros(len(ns), 3, L).pin_memory())
    eng.submit(batch["att_feats"], batch["boxes"], batch["att_masks"], {"beam_size": 3}, slot=1, out=out)
    eng.wait(host=True)
    rseq, rlp = O.sample(z["w"], z["cfg"], batch["att_feats"].clone(), batch["boxes"].clone(), batch["att_masks"].clone(), {"beam_size": 3})
    assert torch.equal(out[0].long(), rseq)
    torch.testing.assert_close(out[1], rlp, rtol=1e-4, atol=2e-5)
    fb16 = FeatureBatcher(str(att_dir), str(box_dir), feat_dim=F, max_boxes=16, max_batch=8, dtype=torch.bfloat16)
    b16 = fb16.from_ids(range(len(ns)))
    eng16 = OrtEngine(z["w"], ModelCfg(z["cfg_dict"]), precision="bf16", device=DEV)
    seq16, _ = eng16.sample(b16["att_feats"], b16["boxes"], b16["att_masks"], {"beam_size": 3})
    assert float((seq16.cpu()[:, 0, 0] == rseq[:, 0, 0]).float().mean()) >= 0.5 and tuple(seq16.shape) == tuple(rseq.shape)


def test_eval_caller_with_radix_detokenisation(tmp_path):
    """eval_on_split's flow (utils/training.py:257-327) on an ACORT-style radix model: batches -> mode="sample" -> best beam ->
    radix digits -> word ids (vectorised, on the device) -> strings -> COCO result JSON."""
    import sparse_caption_b200.relation_transformer as R
    from sparse_caption_b200 import evaluate as E
    z = golden_io.load("acort_tiny")
    m = R.get_model("relation_transformer")(z["cfg_dict"])
    m.load_state_dict(z["w"], strict=True)
    m = m.to(DEV).eval()
    m.precision = "fp32"
    base, tpw = 8, 2                        # radix vocabulary of 8 digits + specials inside the fixture's 35 tokens
    eos = base + 2
    vocab = {i: f"w{i}" for i in range(4 + base ** tpw)}
    dec = lambda ids: " ".join(vocab[i] for i in ids)
    batches = [{"att_feats": z["att_feats"].to(DEV), "boxes": z["boxes"].to(DEV), "image_ids": [7, 9]}]
    preds, speed, path = E.eval_on_split(m, batches, {"beam_size": 3}, decode_words=dec, radix=(base, tpw), vocab_len=len(vocab),
                                         json_fpath=str(tmp_path / "test_beam_3" / "caption_00000000.json"), eos_id=eos)
    assert len(preds) == 2 and speed > 0
    got = json.load(open(path))
    assert [g["image_id"] for g in got] == [7, 9]
    # same captions from the reference's golden token ids through a per-caption restatement of _decode_radix_ids
    for k, cap in enumerate(preds):
        ids = z["beam3_seq"][k, 0].tolist()
        if eos in ids:
            ids = ids[: ids.index(eos)]
        ids = ids + [1] * (-len(ids) % tpw)
        words = [sum(max(d - 1, 0) * base ** i for i, d in enumerate(reversed(ids[j: j + tpw]))) + 4 for j in range(0, len(ids), tpw)]
        words = [w if w < len(vocab) else 1 for w in words]
        assert cap == dec(words), (k, cap)
