"""Parity of the training-side kernels (K2, K4/K9 backward, K10) against torch autograd on the CPU oracle formulas."""
import math

import pytest
import torch

from oracle import ort_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def K():
    import sparse_caption_b200.kernels as k
    return k


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _masked_weight_autograd(W, S, U, mode, bypass):
    """Reference straight-through masked weight (sampler.py:10-66, masked_layer.py:84-110)."""
    if mode == 0:
        return W
    if mode == 3:
        return W * S
    p = torch.sigmoid(S)
    hard = (U < p).to(W.dtype) if mode == 4 else torch.round(p)
    if bypass:
        m = S + (hard - S).detach()
    else:
        m = p + (hard - p).detach()
    return W * m


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mode,bypass", [(0, False), (4, False), (4, True), (3, False), (1, False)])
@pytest.mark.parametrize("shape", [(70, 40, 96), (300, 512, 520), (3000, 136, 128)])  # last one: split-K wgrad
def test_masked_linear_backward(K, dt, mode, bypass, shape):
    """dX via (W.m)^T operand, dW/dS via the fused wgrad epilogue, db via colsum."""
    M, N, Kd = shape
    g = torch.Generator().manual_seed(1)
    x = torch.randn(M, Kd, generator=g)
    W = torch.randn(N, Kd, generator=g) / math.sqrt(Kd)
    S = torch.randn(N, Kd, generator=g)
    U = torch.rand(N, Kd, generator=g)
    dy = torch.randn(M, N, generator=g)
    xq, dyq = x.to(dt).double(), dy.to(dt).double()
    Wd, Sd = W.double().requires_grad_(True), S.double().requires_grad_(True)
    Wm = _masked_weight_autograd(Wd, Sd, U.double(), mode, bypass)
    # operands are quantised to dt on the GPU: mirror that for the reference products
    Wm_q = Wm + (Wm.detach().to(dt).double() - Wm.detach())
    y = xq @ Wm_q.t()
    y.backward(dyq)
    dx_ref = dyq @ Wm.detach().to(dt).double()
    sp = 0.37
    ds_ref = None if mode == 0 else Sd.grad + (sp * torch.sigmoid(S.double()) * (1 - torch.sigmoid(S.double())) if mode in (1, 4) else 0)

    Mp = K.pad8(M)
    dyT = torch.zeros(N, Mp, dtype=dt, device=DEV)
    dyb = torch.zeros(M, N, dtype=dt, device=DEV)
    db_fused = torch.zeros(N, device=DEV)
    K.prep_grad(dy.to(DEV), out=dyb, outT=dyT, colsum=db_fused)
    assert rel_err(db_fused, dyq.sum(0)) < 1e-5  # bias gradient from the same pass
    xT = torch.zeros(Kd, Mp, dtype=dt, device=DEV)
    K.transpose(x.to(dt).to(DEV), xT)
    WmT = torch.zeros(Kd, N, dtype=dt, device=DEV)
    Wg, Sg, Ug = W.to(DEV), S.to(DEV), U.to(DEV)
    Wm_plain = torch.zeros(N, Kd, dtype=dt, device=DEV)
    K.apply_mask_transposed(Wg, Sg if mode else None, mode, WmT, uniforms=Ug, out=Wm_plain)
    assert torch.equal(Wm_plain.t().contiguous(), WmT)  # forward operand and dX operand share the mask sample
    dx = K.linear(dyb, WmT)
    tol = 2e-5 if dt == torch.float32 else 2e-5  # references use the same quantised operands
    assert rel_err(dx, dx_ref) < tol
    dw = torch.zeros(N, Kd, device=DEV)
    ds = torch.zeros(N, Kd, device=DEV)
    K.linear_wgrad(dyT, xT, Wg, Sg if mode else None, mode, dw, ds if mode else None, M=Mp, uniforms=Ug, bypass=bypass,
                   sp_coeff=sp if mode in (1, 4) else 0.0)
    assert rel_err(dw, Wd.grad) < 3e-5
    if mode:
        assert rel_err(ds, ds_ref) < 3e-5
    if dt == torch.bfloat16:
        # two-kernel variant: split-K partial products in a workspace + sc_mask_grad_reduce
        for nsplit in (1, 6):
            wsp = torch.empty(nsplit * N * Kd, device=DEV)
            dw2 = torch.full((N, Kd), 7.0, device=DEV)
            ds2 = torch.full((N, Kd), 7.0, device=DEV)
            K.linear_wgrad(dyT, xT, Wg, Sg if mode else None, mode, dw2, ds2 if mode else None, M=Mp, uniforms=Ug, bypass=bypass,
                           sp_coeff=sp if mode in (1, 4) else 0.0, workspace=wsp)
            assert rel_err(dw2, Wd.grad) < 3e-5
            if mode:
                assert rel_err(ds2, ds_ref) < 3e-5
        if N % 8 == 0 and Kd % 8 == 0:
            # MN-major variant: the row-major activations themselves, no transposed copies, unpadded token count
            wsp = torch.empty(4 * N * Kd, device=DEV)
            dw3 = torch.full((N, Kd), 7.0, device=DEV)
            ds3 = torch.full((N, Kd), 7.0, device=DEV)
            K.linear_wgrad_rowmajor(dyb, x.to(dt).to(DEV), Wg, Sg if mode else None, mode, dw3, ds3 if mode else None, workspace=wsp,
                                    uniforms=Ug, bypass=bypass, sp_coeff=sp if mode in (1, 4) else 0.0)
            assert rel_err(dw3, Wd.grad) < 3e-5
            if mode:
                assert rel_err(ds3, ds_ref) < 3e-5
    db = torch.zeros(N, device=DEV)
    K.colsum(dyb, db)
    assert rel_err(db, dyq.sum(0)) < 1e-5
    # accumulate flag
    K.linear_wgrad(dyT, xT, Wg, Sg if mode else None, mode, dw, None, M=Mp, uniforms=Ug, bypass=bypass, accumulate=True)
    assert rel_err(dw, 2 * Wd.grad) < 3e-5


def test_prep_grad_relu_and_dropout(K):
    g = torch.Generator().manual_seed(2)
    G = torch.randn(100, 72, generator=g)
    H = torch.relu(torch.randn(100, 72, generator=g))
    out = torch.zeros(100, 72, device=DEV, dtype=torch.bfloat16)
    outT = torch.zeros(72, 104, device=DEV, dtype=torch.bfloat16)
    K.prep_grad(G.to(DEV), h=H.to(DEV), out=out, outT=outT, scale=1.25)
    ref = (G * (H != 0) * 1.25).bfloat16()
    assert torch.equal(out.cpu(), ref)
    assert torch.equal(outT.cpu()[:, :100], ref.t())
    assert float(outT[:, 100:].abs().sum()) == 0
    # dropout regeneration matches the forward epilogue's mask
    x = torch.randn(64, 32, generator=g)
    W = torch.eye(32)
    y = K.linear_dropout(x.to(DEV), W.to(DEV), None, p=0.3, drop_seed=11, drop_stream=5)
    keep = (y != 0).float().cpu()
    assert abs(float(keep.mean()) - 0.7) < 0.05
    assert rel_err(y, x * keep / 0.7) < 1e-6
    o2 = torch.zeros(64, 32, device=DEV)
    K.prep_grad(torch.ones(64, 32, device=DEV), out=o2, p=0.3, seed=11, stream_id=5)
    assert rel_err(o2, keep / 0.7) < 1e-6


@pytest.mark.parametrize("D", [64, 512])
def test_layernorm_bwd(K, D):
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(97, D, generator=g) * 2 + 0.5)
    a, b = torch.randn(D, generator=g), torch.randn(D, generator=g)
    dy = torch.randn(97, D, generator=g)
    dres = torch.randn(97, D, generator=g)
    xd, ad, bd = x.double().requires_grad_(True), a.double().requires_grad_(True), b.double().requires_grad_(True)
    O.layer_norm(xd, ad, bd).backward(dy.double())
    dx = torch.zeros(97, D, device=DEV)
    da = torch.zeros(D, device=DEV)
    db = torch.zeros(D, device=DEV)
    K.layernorm_bwd(x.to(DEV), a.to(DEV), dy.to(DEV), dx, da, db, dres=dres.to(DEV))
    assert rel_err(dx, xd.grad + dres.double()) < 1e-5
    assert rel_err(da, ad.grad) < 1e-5
    assert rel_err(db, bd.grad) < 1e-5


def test_logsoftmax_nll(K):
    g = torch.Generator().manual_seed(4)
    R, V = 37, 1000
    logits = torch.randn(R, V, generator=g) * 3
    tgt = torch.randint(0, V, (R,), generator=g)
    w = (torch.rand(R, generator=g) > 0.3).float()
    ld = logits.double().requires_grad_(True)
    lp = torch.log_softmax(ld, -1)
    loss = -(lp.gather(1, tgt.unsqueeze(1)).squeeze(1) * w.double()).sum() / w.sum()
    loss.backward()
    loss_sum = torch.zeros(1, device=DEV)
    inv = torch.tensor([1.0 / float(w.sum())], device=DEV)
    dl = torch.zeros(R, V, device=DEV)
    lpo = torch.zeros(R, V, device=DEV)
    K.logsoftmax_nll(logits.to(DEV), tgt.int().to(DEV), w.to(DEV), inv, loss_sum, dl, lpo)
    assert rel_err(loss_sum * inv, loss.detach()) < 1e-5
    assert rel_err(dl, ld.grad) < 1e-5
    assert rel_err(lpo, lp.detach()) < 1e-5


def test_embedding_bwd_and_mask_grad(K):
    g = torch.Generator().manual_seed(5)
    V, D, R = 40, 64, 90
    tok = torch.randint(0, V, (R,), generator=g)
    dy = torch.randn(R, D, generator=g)
    W, S, U = torch.randn(V, D, generator=g), torch.randn(V, D, generator=g), torch.rand(V, D, generator=g)
    Wd, Sd = W.double().requires_grad_(True), S.double().requires_grad_(True)
    Wm = _masked_weight_autograd(Wd, Sd, U.double(), 4, False)
    (Wm[tok] * 8.0).backward(dy.double())
    dtab = torch.zeros(V, D, device=DEV)
    K.embedding_bwd(tok.int().to(DEV), dy.to(DEV), dtab, 8.0)
    dw = torch.zeros(V, D, device=DEV)
    ds = torch.zeros(V, D, device=DEV)
    K.mask_grad(dtab, W.to(DEV), S.to(DEV), 4, dw, ds, uniforms=U.to(DEV))
    assert rel_err(dw, Wd.grad) < 1e-5
    assert rel_err(ds, Sd.grad) < 1e-5


def test_adam_clip(K):
    g = torch.Generator().manual_seed(6)
    n = 10007
    p0 = torch.randn(n, generator=g)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=0.01, betas=(0.9, 0.98), eps=1e-9, weight_decay=0.01)
    p = p0.to(DEV).clone()
    m = torch.zeros(n, device=DEV)
    v = torch.zeros(n, device=DEV)
    for step in range(1, 4):
        gr = torch.randn(n, generator=g) * 0.3
        p_ref.grad = gr.clone()
        torch.nn.utils.clip_grad_value_([p_ref], 0.1)
        opt.step()
        K.adam_clip(p, gr.to(DEV), m, v, lr=0.01, betas=(0.9, 0.98), eps=1e-9, weight_decay=0.01, clip=0.1, grad_scale=1.0,
                    step=step)
    assert rel_err(p, p_ref.detach()) < 1e-5


def _attn_ref(q, k, v, h, G, Tq, Tk, key_valid, bias, causal_T):
    D = q.shape[1]
    dk = D // h
    qh = q.view(G, Tq, h, dk).transpose(1, 2)
    kh = k.view(G, Tk, h, dk).transpose(1, 2)
    vh = v.view(G, Tk, h, dk).transpose(1, 2)
    s = qh @ kh.transpose(-2, -1) / math.sqrt(dk)
    mask = torch.ones(G, 1, Tq, Tk, dtype=torch.bool)
    if key_valid is not None:
        mask = mask & (key_valid.view(G, 1, 1, Tk) != 0)
    if causal_T:
        i = torch.arange(Tq).view(1, 1, Tq, 1) % causal_T
        j = torch.arange(Tk).view(1, 1, 1, Tk)
        mask = mask & (j <= i)
    s = s.masked_fill(~mask, -1e9)
    if bias is not None:
        s = s + bias
    p = torch.softmax(s, -1)
    return (p @ vh).transpose(1, 2).reshape(G * Tq, D), p


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", ["self", "cross", "box"])
def test_attention_fwd_bwd(K, dt, case):
    g = torch.Generator().manual_seed(7)
    h, dk = 4, 16
    D = h * dk
    if case == "self":
        G, Tq, Tk, causal = 6, 9, 9, 9
        kv = (torch.rand(G, Tk, generator=g) > 0.2).float()
        kv[:, 0] = 1
        bias = None
    elif case == "cross":
        G, Tq, Tk, causal = 3, 2 * 9, 11, 0
        kv = torch.ones(G, Tk)
        kv[1, 8:] = 0
        bias = None
    else:
        G, Tq, Tk, causal = 3, 13, 13, 0
        kv = torch.ones(G, Tk)
        kv[2, 10:] = 0
        bias = torch.randn(G, h, Tq, Tk, generator=g)
    q = torch.randn(G * Tq, D, generator=g).to(dt)
    k = torch.randn(G * Tk, D, generator=g).to(dt)
    v = torch.randn(G * Tk, D, generator=g).to(dt)
    dO = torch.randn(G * Tq, D, generator=g)
    qd, kd, vd = (t.double().requires_grad_(True) for t in (q, k, v))
    bd = bias.double().requires_grad_(True) if bias is not None else None
    out_ref, p_ref = _attn_ref(qd, kd, vd, h, G, Tq, Tk, kv, bd, causal)
    out_ref.backward(dO.double())
    out = torch.zeros(G * Tq, D, dtype=dt, device=DEV)
    probs = torch.zeros(G, h, Tq, Tk, device=DEV)
    qg, kg, vg = q.to(DEV), k.to(DEV), v.to(DEV)
    K.attention_fwd(qg, kg, vg, out, probs, G=G, Tq=Tq, Tk=Tk, h=h, dk=dk, ldq=D, ldk=D, ldv=D, ldo=D, key_valid=kv.to(DEV),
                    bias=None if bias is None else bias.to(DEV), causal_T=causal)
    tol = 1e-5 if dt == torch.float32 else 1e-2
    assert rel_err(out.float(), out_ref.detach()) < tol
    assert rel_err(probs, p_ref.detach()) < 1e-5
    dq = torch.zeros(G * Tq, D, device=DEV)
    dkk = torch.zeros(G * Tk, D, device=DEV)
    dv = torch.zeros(G * Tk, D, device=DEV)
    dbias = torch.zeros(G, h, Tq, Tk, device=DEV) if bias is not None else None
    K.attention_bwd(qg, kg, vg, probs, dO.to(DEV), dq, dkk, dv, dtype=dt, G=G, Tq=Tq, Tk=Tk, h=h, dk=dk, ldq=D, ldk=D, ldv=D, ldd=D,
                    ldgq=D, ldgk=D, ldgv=D, dbias=dbias)
    assert rel_err(dq, qd.grad) < 2e-5
    assert rel_err(dkk, kd.grad) < 2e-5
    assert rel_err(dv, vd.grad) < 2e-5
    if bias is not None:
        assert rel_err(dbias, bd.grad) < 2e-5


def test_attention_dropout_consistency(K):
    """The backward regenerates the forward's dropout mask: check d(out)/d(v) against the saved probabilities."""
    g = torch.Generator().manual_seed(8)
    G, T, h, dk = 4, 8, 2, 16
    D = h * dk
    q, k, v = (torch.randn(G * T, D, generator=g).to(DEV) for _ in range(3))
    out = torch.zeros(G * T, D, device=DEV)
    probs = torch.zeros(G, h, T, T, device=DEV)
    K.attention_fwd(q, k, v, out, probs, G=G, Tq=T, Tk=T, h=h, dk=dk, ldq=D, ldk=D, ldv=D, ldo=D, p=0.4, seed=3, stream_id=9)
    # recover the effective (dropped, rescaled) probabilities from out with v = identity-like probe
    dO = torch.randn(G * T, D, generator=g).to(DEV)
    dq, dkk, dv = (torch.zeros(G * T, D, device=DEV) for _ in range(3))
    K.attention_bwd(q, k, v, probs, dO, dq, dkk, dv, dtype=torch.float32, G=G, Tq=T, Tk=T, h=h, dk=dk, ldq=D, ldk=D, ldv=D, ldd=D,
                    ldgq=D, ldgk=D, ldgv=D, p=0.4, seed=3, stream_id=9)
    # dV = Pd^T dO and out = Pd V  =>  <out, dO> == <V, dV>
    lhs = float((out * dO).sum())
    rhs = float((v * dv).sum())
    assert abs(lhs - rhs) < 1e-3 * max(1.0, abs(lhs))


def test_box_bias_fwd_bwd(K):
    g = torch.Generator().manual_seed(9)
    B, N, h = 3, 12, 8
    boxes = O.synthetic_inputs(B, N, 8, seed=5)["boxes"]
    wg_w = (torch.randn(h, 64, generator=g) * 0.3)
    wg_b = (torch.randn(h, generator=g) * 0.3)
    emb = O.box_relational_embedding(boxes).double()
    ww, wb = wg_w.double().requires_grad_(True), wg_b.double().requires_grad_(True)
    gg = torch.relu(torch.einsum("bijf,hf->bhij", emb, ww) + wb.view(1, h, 1, 1))
    bias_ref = torch.log(torch.clamp(gg, min=1e-6))
    dbias = torch.randn(B, h, N, N, generator=g)
    bias_ref.backward(dbias.double())
    bias = torch.zeros(B, h, N, N, device=DEV)
    K.box_bias_fwd(boxes.to(DEV), wg_w.to(DEV), wg_b.to(DEV), bias, B=B, N=N, h=h)
    active = bias_ref.detach() > -13
    assert rel_err(bias[active.to(DEV)], bias_ref.detach()[active]) < 5e-4
    dw = torch.zeros(h, 64, device=DEV)
    db = torch.zeros(h, device=DEV)
    K.box_bias_bwd(boxes.to(DEV), bias, dbias.to(DEV), dw, db, B=B, N=N, h=h)
    assert rel_err(dw, ww.grad) < 2e-3
    assert rel_err(db, wb.grad) < 2e-3


@pytest.mark.parametrize("case", ["self", "cross", "box", "wide"])
def test_attention_fwd_bwd_tensor_path(K, case):
    """bf16, d_k = 64: sc_attention_fwd / _bwd run the mma.sync kernels (sc_mma_attention_train.cu).  Operands are
    bf16-exact; P, dS and dO are rounded to bf16 inside the kernel, hence the 2e-2 bound north_star states for bf16."""
    g = torch.Generator().manual_seed(11)
    h, dk = 2, 64
    D = h * dk
    if case == "self":
        G, Tq, Tk, causal = 5, 17, 17, 17
        kv = (torch.rand(G, Tk, generator=g) > 0.2).float()
        kv[:, 0] = 1
        bias = None
    elif case == "cross":
        G, Tq, Tk, causal = 3, 5 * 17, 36, 0
        kv = torch.ones(G, Tk)
        kv[1, 30:] = 0
        bias = None
    elif case == "box":
        G, Tq, Tk, causal = 3, 36, 36, 0
        kv = torch.ones(G, Tk)
        kv[2, 33:] = 0
        bias = torch.randn(G, h, Tq, Tk, generator=g)
    else:
        G, Tq, Tk, causal = 2, 150, 50, 0   # more m-tiles than warps, 4 key tiles
        kv = None
        bias = torch.randn(G, h, Tq, Tk, generator=g)
    dt = torch.bfloat16
    q = torch.randn(G * Tq, D, generator=g).to(dt)
    k = torch.randn(G * Tk, D, generator=g).to(dt)
    v = torch.randn(G * Tk, D, generator=g).to(dt)
    dO = torch.randn(G * Tq, D, generator=g)
    qd, kd, vd = (t.double().requires_grad_(True) for t in (q, k, v))
    bd = bias.double().requires_grad_(True) if bias is not None else None
    out_ref, p_ref = _attn_ref(qd, kd, vd, h, G, Tq, Tk, kv, bd, causal)
    out_ref.backward(dO.double())
    out = torch.zeros(G * Tq, D, dtype=dt, device=DEV)
    probs = torch.zeros(G, h, Tq, Tk, device=DEV)
    qg, kg, vg = q.to(DEV), k.to(DEV), v.to(DEV)
    K.attention_fwd(qg, kg, vg, out, probs, G=G, Tq=Tq, Tk=Tk, h=h, dk=dk, ldq=D, ldk=D, ldv=D, ldo=D,
                    key_valid=None if kv is None else kv.to(DEV), bias=None if bias is None else bias.to(DEV), causal_T=causal)
    assert rel_err(out.float(), out_ref.detach()) < 2e-2
    assert rel_err(probs, p_ref.detach()) < 1e-4  # scores accumulate in fp32 from bf16-exact operands
    dq = torch.full((G * Tq, D), float("nan"), device=DEV)
    dkk = torch.full((G * Tk, D), float("nan"), device=DEV)
    dv = torch.full((G * Tk, D), float("nan"), device=DEV)
    dbias = torch.zeros(G, h, Tq, Tk, device=DEV) if bias is not None else None
    K.attention_bwd(qg, kg, vg, probs, dO.to(DEV), dq, dkk, dv, dtype=dt, G=G, Tq=Tq, Tk=Tk, h=h, dk=dk, ldq=D, ldk=D, ldv=D, ldd=D,
                    ldgq=D, ldgk=D, ldgv=D, dbias=dbias)
    assert rel_err(dq, qd.grad) < 2e-2
    assert rel_err(dkk, kd.grad) < 2e-2
    assert rel_err(dv, vd.grad) < 2e-2
    if bias is not None:
        assert rel_err(dbias, bd.grad) < 2e-2


def test_attention_dropout_consistency_tensor_path(K):
    """bf16 / d_k = 64: the backward regenerates the forward's dropout mask (<out, dO> == <V, dV>)."""
    g = torch.Generator().manual_seed(12)
    G, Tq, Tk, h, dk = 4, 40, 36, 2, 64
    D = h * dk
    q = torch.randn(G * Tq, D, generator=g).to(DEV).bfloat16()
    k, v = (torch.randn(G * Tk, D, generator=g).to(DEV).bfloat16() for _ in range(2))
    out = torch.zeros(G * Tq, D, device=DEV, dtype=torch.bfloat16)
    probs = torch.zeros(G, h, Tq, Tk, device=DEV)
    K.attention_fwd(q, k, v, out, probs, G=G, Tq=Tq, Tk=Tk, h=h, dk=dk, ldq=D, ldk=D, ldv=D, ldo=D, p=0.4, seed=3, stream_id=9)
    assert abs(float(probs.sum()) - G * h * Tq) < 1e-2  # saved probabilities are pre-dropout
    dO = torch.randn(G * Tq, D, generator=g).to(DEV)
    dq, dkk, dv = torch.zeros(G * Tq, D, device=DEV), torch.zeros(G * Tk, D, device=DEV), torch.zeros(G * Tk, D, device=DEV)
    K.attention_bwd(q, k, v, probs, dO, dq, dkk, dv, dtype=torch.bfloat16, G=G, Tq=Tq, Tk=Tk, h=h, dk=dk, ldq=D, ldk=D, ldv=D, ldd=D,
                    ldgq=D, ldgk=D, ldgv=D, p=0.4, seed=3, stream_id=9)
    lhs = float((out.float() * dO).sum())
    rhs = float((v.float() * dv).sum())
    assert abs(lhs - rhs) < 2e-2 * max(1.0, abs(lhs))
    # the same mask, element for element: out equals (probs (.) keep) V with keep recovered from the fp32 kernel's rule
    out32 = torch.zeros(G * Tq, D, device=DEV)
    probs32 = torch.zeros(G, h, Tq, Tk, device=DEV)
    K.attention_fwd(q.float(), k.float(), v.float(), out32, probs32, G=G, Tq=Tq, Tk=Tk, h=h, dk=dk, ldq=D, ldk=D, ldv=D, ldo=D,
                    p=0.4, seed=3, stream_id=9)
    assert rel_err(out.float(), out32) < 2e-2


def test_apply_mask_batched_matches_per_tensor(K):
    """One launch for all masked weights == the per-tensor kernel, bit for bit (same Philox sample), in every mode."""
    g = torch.Generator().manual_seed(13)
    shapes = [(512, 512), (200, 72), (1000, 512), (64, 2048), (8, 64)]
    for mode in (K.MASK_BERNOULLI, K.MASK_ROUND, K.MASK_UNIFORM, K.MASK_RAW, K.MASK_NONE):
        items, refs = [], []
        for i, (N, Kd) in enumerate(shapes):
            W = torch.randn(N, Kd, generator=g).to(DEV)
            S = (torch.randn(N, Kd, generator=g) * 2).to(DEV)
            U = torch.rand(N, Kd, generator=g).to(DEV)
            out = torch.full((N, Kd), 7.0, device=DEV, dtype=torch.bfloat16)
            outT = torch.full((Kd, N), 7.0, device=DEV, dtype=torch.bfloat16)
            items.append((W, None if mode == K.MASK_NONE else S, U if mode == K.MASK_UNIFORM else None, out, outT, 5 + i))
            r, rT = torch.empty_like(out), torch.empty_like(outT)
            K.apply_mask_transposed(W, None if mode == K.MASK_NONE else S, mode, rT, uniforms=U if mode == K.MASK_UNIFORM else None,
                                    seed=99, stream_id=4096 + 5 + i, out=r)
            refs.append((r, rT))
        desc, tiles = K.mask_descriptors(items, DEV)
        K.apply_mask_batched(desc, tiles, mode, seed=99, stream_base=4096)
        for (W, S, U, out, outT, sid), (r, rT) in zip(items, refs):
            assert torch.equal(out, r), (mode, tuple(W.shape))
            assert torch.equal(outT, rT), (mode, tuple(W.shape))


@pytest.mark.parametrize("shape", [(4250, 2048, 512), (300, 520, 136), (70, 96, 40)])
def test_linear_hmask(K, shape):
    """dX GEMM whose epilogue applies the saved-activation mask, stores bf16 and accumulates the bias gradient."""
    M, N, Kd = shape
    g = torch.Generator().manual_seed(41)
    x = torch.randn(M, Kd, generator=g).bfloat16().to(DEV)
    w = (torch.randn(N, Kd, generator=g) * 0.1).bfloat16().to(DEV)
    h = (torch.randn(M, N, generator=g).clamp_min(0) * (torch.rand(M, N, generator=g) > 0.3)).bfloat16().to(DEV)
    out = torch.full((M, N), 9.0, device=DEV, dtype=torch.bfloat16)
    colsum = torch.full((N,), 0.5, device=DEV)
    K.linear_hmask(x, w, h, out, scale=1.25, colsum=colsum)
    ref = (x.double() @ w.double().t()) * 1.25 * (h.double() != 0)
    assert rel_err(out.float(), ref) < 1e-2
    assert bool((out[h == 0] == 0).all())
    assert rel_err(colsum - 0.5, out.double().sum(0)) < 1e-4   # sums of the bf16 values that were stored
