"""Drop-in class surface on the GPU: the model registry classes and the masked layers behave like the reference's
(golden fixtures / oracle), through the kernels only."""
import pytest
import torch

from oracle import ort_oracle as O
from tests import golden_io

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_dense_model_class_matches_reference():
    import sparse_caption_b200.relation_transformer as R
    z = golden_io.load("ort_tiny")
    m = R.get_model("relation_transformer")(z["cfg_dict"])
    m.load_state_dict(z["w"], strict=True)
    m = m.to(DEV).eval()
    m.precision = "fp32"
    seq, lp = m(att_feats=z["att_feats"].to(DEV), boxes=z["boxes"].to(DEV), opt={"beam_size": 3}, mode="sample")
    assert seq.dtype == torch.int64 and torch.equal(seq.cpu(), z["beam3_seq"])
    torch.testing.assert_close(lp.cpu(), z["beam3_lp"], rtol=1e-4, atol=2e-5)
    out = m(att_feats=z["att_feats"].to(DEV), boxes=z["boxes"].to(DEV), seqs=z["seqs"].to(DEV))
    assert out.shape == z["tf_logprobs"].shape
    assert rel_err(out, z["tf_logprobs"]) < 2e-4


def test_prune_model_class_eval_and_training_step():
    import sparse_caption_b200.relation_transformer as R
    z = golden_io.load("ort_prune_tiny")
    m = R.get_model("relation_transformer_prune")(z["cfg_dict"])
    m.load_state_dict(z["w"], strict=True)
    m = m.to(DEV).eval()
    m.precision = "fp32"
    out = m(att_feats=z["att_feats"].to(DEV), boxes=z["boxes"].to(DEV), seqs=z["seqs"].to(DEV))
    assert rel_err(out, z["tf_logprobs_eval"]) < 2e-4
    # sampling with binarized masks == dense class on folded weights (eval_model.py:64-77 flow)
    seq, lp = m(att_feats=z["att_feats"].to(DEV), boxes=z["boxes"].to(DEV), opt={"beam_size": 2}, mode="sample")
    eff = O.effective_state_dict(z["w"], "supermask")
    rseq, rlp = O.sample(eff, z["cfg"], z["att_feats"], z["boxes"], None, {"beam_size": 2})
    assert torch.equal(seq.cpu(), rseq)
    torch.testing.assert_close(lp.cpu(), rlp, rtol=1e-4, atol=2e-5)
    # one kernel-side training step, then parameters flow back into the module
    m.train()
    m._trainer = None
    tr = m.trainer(seed=3)
    before = m.model.generator.proj.weight.detach().clone()
    loss = tr.train_step(z["att_feats"].to(DEV), z["boxes"].to(DEV), z["seqs"], z["masks"], seq_per_img=2, lr=1e-3,
                         sparsity_target=0.9, sparsity_weight=5.0, current_step=5, max_step=10)
    assert torch.isfinite(loss).all()
    m.sync_from_trainer()
    assert not torch.equal(before, m.model.generator.proj.weight.detach())


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_masked_linear_layer_autograd(precision):
    from sparse_caption_b200 import masked_layer as ML
    ML.set_precision(precision)
    try:
        g = torch.Generator().manual_seed(0)
        lin = ML.MaskedLinear(64, 48, "supermask", 5.0).to(DEV)
        with torch.no_grad():
            lin.weight_pruning_mask.copy_(torch.randn(48, 64, generator=g) * 2)
        x = torch.randn(10, 7, 64, generator=g).to(DEV).requires_grad_(True)
        lin.eval()
        y = lin(x)
        Wm = lin.weight.detach().cpu() * O.binarize_logits(lin.weight_pruning_mask.detach().cpu())
        xr = x.detach().cpu()
        if precision == "bf16":
            xr, Wm = xr.bfloat16().float(), Wm.bfloat16().float()
        ref = torch.nn.functional.linear(xr, Wm, lin.bias.detach().cpu())
        assert rel_err(y, ref) < (1e-5 if precision == "fp32" else 1e-2)
        gy = torch.randn(y.shape, generator=g).to(DEV)
        y.backward(gy)
        # straight-through: dS = (dy^T x) * W * sigmoid'(S); dW = (dy^T x) * m
        dWm = gy.cpu().reshape(-1, 48).t() @ x.detach().cpu().reshape(-1, 64)
        S = lin.weight_pruning_mask.detach().cpu()
        sig = torch.sigmoid(S)
        tol = 1e-4 if precision == "fp32" else 3e-2
        assert rel_err(lin.weight.grad, dWm * O.binarize_logits(S)) < tol
        assert rel_err(lin.weight_pruning_mask.grad, dWm * lin.weight.detach().cpu() * sig * (1 - sig)) < tol
        assert rel_err(lin.bias.grad, gy.cpu().reshape(-1, 48).sum(0)) < tol
        assert rel_err(x.grad, (gy.cpu().reshape(-1, 48) @ Wm).reshape(x.shape)) < tol
    finally:
        ML.set_precision("bf16")


def test_supermask_toy_training_reaches_sparsity():
    """Mirrors the reference's tests/test_prune.py:119-125: 40 Adam steps with the sparsity loss on a toy model of
    masked layers drive the binarized-mask sparsity to the target within 0.3."""
    from sparse_caption_b200 import masked_layer as ML, sampler
    from sparse_caption_b200.prune import PruningMixin
    ML.set_precision("fp32")
    sampler.set_mask_seed(8888)

    class Toy(PruningMixin, torch.nn.Module):
        def __init__(self):
            super().__init__(mask_type="supermask", mask_freeze_scope="")
            self.emb = ML.MaskedEmbedding(20, 32, "supermask", 5.0)
            self.l1 = ML.MaskedLinear(32, 64, "supermask", 5.0)
            self.l2 = ML.MaskedLinear(64, 8, "supermask", 5.0)

        def forward(self, ids):
            return self.l2(torch.relu(self.l1(self.emb(ids))))

    try:
        torch.manual_seed(8888)
        m = Toy().to(DEV)
        assert float(m.all_mask_sparsities[0]) == 0.0
        opt = torch.optim.Adam([{"params": m.all_weights(named=False), "lr": 1e-2},
                                {"params": m.all_pruning_masks(named=False), "lr": 100.0, "eps": 1e-2}])
        ids = torch.randint(0, 20, (16, 5), device=DEV)
        tgt = torch.randn(16, 5, 8, device=DEV)
        for step in range(40):
            opt.zero_grad()
            loss = ((m(ids) - tgt) ** 2).mean() + m.compute_sparsity_loss(0.8, 120.0, step, 40)
            loss.backward()
            opt.step()
        sp = float(m.all_mask_sparsities[0])
        assert abs(sp - 0.8) < 0.3, sp
        m.prune_weights()
        assert abs(float(m.all_weight_sparsities[0]) - sp) < 0.05
    finally:
        ML.set_precision("bf16")
