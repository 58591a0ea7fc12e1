"""CIDEr-D reward: the oracle against the IMPORTED reference scorer's outputs (golden fixture, CPU tier),
the device kernel against the oracle (GPU tier), and the sample / baseline bookkeeping of CaptionScorer."""
import os

import numpy as np
import pytest
import torch

from oracle import cider_oracle as C

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ciderd.npz")


def _load():
    z = np.load(GOLDEN)
    unpad = lambda a: [[int(x) for x in r if x >= 0] for r in a]
    refs_flat, hyps = unpad(z["refs"]), unpad(z["hyps"])
    refs, o = [], 0
    for c in z["ref_count"]:
        refs.append(refs_flat[o: o + int(c)])
        o += int(c)
    ns = len(hyps) // len(refs)
    table = {tuple(int(x) for x in k if x >= 0): float(c) for k, c in zip(z["df_ngrams"], z["df_counts"])}
    return z, refs, hyps, ns, table


def test_oracle_matches_reference_scorer_golden():
    """Bit-exact: same Python-float arithmetic in the same order as ciderD_scorer.py:133-212, both df modes."""
    z, refs, hyps, ns, table = _load()
    rr = [refs[i // ns] for i in range(len(hyps))]
    df = C.corpus_document_frequency(rr)  # the scorer appends an image's references once per hypothesis
    got = C.ciderd(hyps, rr, df, np.log(float(len(rr))))
    assert np.array_equal(got, z["corpus_scores"])
    got = C.ciderd(hyps, rr, table, np.log(float(z["df_docs"])))
    assert np.array_equal(got, z["cached_scores"])
    assert float(got.max()) > 1.0 and float(got.min()) == 0.0  # copies of references score high, empty captions 0


def test_caption_scorer_bookkeeping():
    z, refs, hyps, ns, table = _load()
    B = len(refs)
    sample = [hyps[i * ns: (i + 1) * ns] for i in range(B)]
    ref_len = np.log(float(z["df_docs"]))
    s, b = C.caption_scorer(refs, sample, None, table, ref_len)
    assert s.shape == b.shape == (B * ns,)
    tot = s.reshape(B, ns).sum(-1)
    np.testing.assert_allclose(b.reshape(B, ns), (tot[:, None] - s.reshape(B, ns)) / (ns - 1))
    base = [[refs[i][0]] for i in range(B)]
    s2, b2 = C.caption_scorer(refs, sample, base, table, ref_len)
    assert np.array_equal(s2, s)
    assert np.array_equal(b2.reshape(B, ns), np.repeat(C.ciderd([x[0] for x in base], refs, table, ref_len)[:, None], ns, 1))


@pytest.mark.gpu
def test_device_ciderd_matches_oracle_and_reference():
    from sparse_caption_b200.cider import CiderD
    z, refs, hyps, ns, table = _load()
    B, L = len(refs), 16
    H = len(hyps)
    tok = torch.zeros(H, L, dtype=torch.int32)
    for i, h in enumerate(hyps):                      # <eos> = 3 after the caption, garbage behind it on every other row
        tok[i, : len(h)] = torch.tensor(h, dtype=torch.int32)
        if len(h) < L:
            tok[i, len(h)] = 3
            if i % 2 and len(h) + 1 < L:
                tok[i, len(h) + 1:] = 7
    img = torch.arange(B).repeat_interleave(ns)
    sc = CiderD(table, float(z["df_docs"]), device="cuda")
    sc.set_refs(refs)
    got = sc.score(tok.cuda(), img.cuda()).cpu().numpy()
    np.testing.assert_allclose(got, z["cached_scores"], rtol=1e-12, atol=1e-13)   # the imported reference's scores
    # SCST reward bookkeeping (scorers.py:100-106), with and without a greedy baseline
    s, b = sc.scst_reward(tok.view(B, ns, L).cuda())
    os_, ob = C.caption_scorer(refs, [hyps[i * ns: (i + 1) * ns] for i in range(B)], None, table, np.log(float(z["df_docs"])))
    np.testing.assert_allclose(s.cpu().numpy(), os_, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(b.cpu().numpy(), ob, rtol=1e-11, atol=1e-12)
    greedy = tok.view(B, ns, L)[:, :1].cuda()
    s2, b2 = sc.scst_reward(tok.view(B, ns, L).cuda(), greedy)
    np.testing.assert_allclose(b2.cpu().numpy(), np.repeat(z["cached_scores"].reshape(B, ns)[:, 0], ns), rtol=1e-12, atol=1e-13)
    # "corpus" mode built from the references themselves
    rr = [refs[i // ns] for i in range(H)]
    df = C.corpus_document_frequency(refs)
    want = C.ciderd(hyps, rr, df, np.log(float(B)))
    sc2 = CiderD.from_corpus(refs, device="cuda")
    np.testing.assert_allclose(sc2.score(tok.cuda(), img.cuda()).cpu().numpy(), want, rtol=1e-12, atol=1e-13)


@pytest.mark.gpu
def test_full_scst_step_stays_on_device():
    """rollout (train-mode Bernoulli masks) -> greedy baseline -> CIDEr-D reward on the device -> RewardCriterion on the
    teacher-forced log-probs of the sampled captions -> backward through the module path: the reference's compute_scst_loss
    (utils/training.py:202-255) without a host round trip of the captions."""
    import sparse_caption_b200.relation_transformer as R
    from sparse_caption_b200.cider import CiderD
    from tests import golden_io
    z = golden_io.load("ort_prune_tiny")
    m = R.get_model("relation_transformer_prune")(z["cfg_dict"])
    m.load_state_dict(z["w"], strict=True)
    m = m.cuda()
    m.precision = "fp32"
    att, boxes = z["att_feats"].cuda(), z["boxes"].cuda()
    B = att.shape[0]
    refs = [[[int(t) for t in row[1:] if t not in (0, 3)] for row in z["seqs"][i * 2: (i + 1) * 2]] for i in range(B)]
    scorer = CiderD.from_corpus(refs, device="cuda")
    m.eval()
    greedy, _ = m(att_feats=att, boxes=boxes, opt={"beam_size": 1}, mode="sample")
    m.train()
    sample, _ = m(att_feats=att, boxes=boxes, opt={"num_random_sample": 3, "beam_size": 0, "sample_seed": 2}, mode="sample")
    sc_s, sc_b = scorer.scst_reward(sample, greedy)
    reward = (sc_s - sc_b).float()
    assert reward.is_cuda and tuple(reward.shape) == (B * 3,) and torch.isfinite(reward).all()
    # teacher-forced log-probs of the sampled captions (differentiable), RewardCriterion (utils/losses.py:10-29)
    L = sample.shape[-1]
    seqs = torch.cat([torch.full((B * 3, 1), m.bos_idx, device="cuda"), sample.view(B * 3, L)], 1)
    logp = m(att_feats=att, boxes=boxes, seqs=torch.cat([seqs, torch.zeros(B * 3, 1, dtype=torch.long, device="cuda")], 1))
    lp = logp[:, :L].gather(2, sample.view(B * 3, L, 1)).squeeze(2)
    mask = (sample.view(B * 3, L) != m.pad_idx).float()
    mask = torch.cat([mask.new_ones(B * 3, 1), mask[:, :-1]], 1)
    loss = (-lp * reward[:, None] * mask).sum() / mask.sum()
    loss.backward()
    grads = [p.grad for p in m.parameters() if p.requires_grad]
    assert all(g is not None and torch.isfinite(g).all() for g in grads)
    assert float(reward.abs().max()) > 0 and sum(float(g.abs().sum()) for g in grads) > 0
