"""Pin the CPU oracle (oracle/ort_oracle.py) against outputs of the imported reference
(tests/golden/*.npz).  CPU-only; runs in the `-m "not gpu"` tier."""
import pytest
import torch

from oracle import ort_oracle as O
from tests import golden_io

TOL = dict(rtol=1e-5, atol=2e-6)


def test_binarize_bit_exact():
    z = golden_io.load("binarize")
    assert torch.equal(O.binarize_logits(z["s"]), z["m"])


def test_box_embedding_and_attention():
    z = golden_io.load("box_geometry")
    emb = O.box_relational_embedding(z["boxes"])
    assert torch.equal(emb, z["emb"])  # same torch ops in the same order -> bit exact
    scores = torch.matmul(z["q"], z["k"].transpose(-2, -1)) / (z["q"].size(-1) ** 0.5)
    scores = scores.masked_fill(z["mask"] == 0, -1e9)
    w = torch.softmax(torch.log(torch.clamp(z["g"], min=1e-6)) + scores, -1)
    torch.testing.assert_close(w, z["w"], **TOL)
    torch.testing.assert_close(torch.matmul(w, z["v"]), z["out"], **TOL)


@pytest.mark.parametrize("name", ["ort_tiny", "ort_tiny_masks", "acort_tiny"])
def test_dense_class(name):
    z = golden_io.load(name)
    cfg, sd = z["cfg"], z["w"]
    am = z.get("att_masks")
    lp = O.forward_tf(sd, cfg, z["att_feats"], z["boxes"], z["seqs"], am)
    torch.testing.assert_close(lp, z["tf_logprobs"], **TOL)
    loss = O.lm_criterion(lp, z["seqs"][:, 1:], z["masks"][:, 1:])
    torch.testing.assert_close(loss, z["tf_loss"], **TOL)
    for beam in (3, 2):
        seq, slp = O.sample(sd, cfg, z["att_feats"], z["boxes"], am, {"beam_size": beam})
        assert torch.equal(seq, z[f"beam{beam}_seq"]), name
        torch.testing.assert_close(slp, z[f"beam{beam}_lp"], **TOL)
    seq, slp = O.sample(sd, cfg, z["att_feats"], z["boxes"], am,
                        {"beam_size": 3, "decoding_constraint": 1, "length_penalty": "wu_0.5"})
    assert torch.equal(seq, z["beam3c_seq"])
    torch.testing.assert_close(slp, z["beam3c_lp"], **TOL)
    seq, slp = O.sample(sd, cfg, z["att_feats"], z["boxes"], am, {"beam_size": 1})
    assert torch.equal(seq, z["greedy_seq"])
    torch.testing.assert_close(slp, z["greedy_lp"], **TOL)


def test_dense_class_gradients():
    z = golden_io.load("ort_tiny")
    cfg = z["cfg"]
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(".pe")) for k, v in z["w"].items()}
    lp = O.forward_tf(sd, cfg, z["att_feats"], z["boxes"], z["seqs"], None)
    O.lm_criterion(lp, z["seqs"][:, 1:], z["masks"][:, 1:]).backward()
    for k, g in z["g"].items():
        torch.testing.assert_close(sd[k].grad, g, rtol=1e-4, atol=1e-6, msg=k)


def test_prune_class_eval_and_train():
    z = golden_io.load("ort_prune_tiny")
    cfg, full = z["cfg"], z["w"]
    eff = O.effective_state_dict(full, "supermask", training=False)
    lp = O.forward_tf(eff, cfg, z["att_feats"], z["boxes"], z["seqs"], None)
    torch.testing.assert_close(lp, z["tf_logprobs_eval"], **TOL)
    logits = [full[k] for k in full if k.endswith("_pruning_mask")]
    sl, _ = O.sparsity_loss(logits, 0.8, 7.5, 30, 100)
    torch.testing.assert_close(sl.float(), z["sparsity_loss"], **TOL)
    # train mode: injected uniforms, straight-through gradients (sampler.py:10-34)
    W = {k: v.clone().requires_grad_(True) for k, v in full.items() if k in z["u"]}
    S = {k: full[k + "_pruning_mask"].clone().requires_grad_(True) for k in z["u"]}
    eff = dict(full)
    for k in z["u"]:
        p = torch.sigmoid(S[k])
        m = (z["u"][k] < p).float()
        m = p + (m - p).detach()  # straight-through: d m / d p = 1
        eff[k] = m * W[k]
    eff = {k: v for k, v in eff.items() if not k.endswith("_pruning_mask")}
    lp = O.forward_tf(eff, cfg, z["att_feats"], z["boxes"], z["seqs"], None)
    torch.testing.assert_close(lp, z["tf_logprobs_train"], **TOL)
    O.lm_criterion(lp, z["seqs"][:, 1:], z["masks"][:, 1:]).backward()
    for k, g in z["g"].items():
        if k.endswith("_pruning_mask"):
            torch.testing.assert_close(S[k[: -len("_pruning_mask")]].grad, g, rtol=1e-4, atol=1e-7, msg=k)
        else:
            torch.testing.assert_close(W[k].grad, g, rtol=1e-4, atol=1e-7, msg=k)
