"""CPU tier, world_size 2 over gloo: the data-parallel recipe of sparse_caption_b200/distributed.py reproduces the
single-process gradient — images sharded by rank, each rank's loss normalised by the GLOBAL token count, SUM
all-reduce of weight AND mask-logit gradients, same Bernoulli uniforms on every rank.  The per-rank compute is done
with the CPU oracle (this tier has no GPU); the GPU trainer calls the same helpers."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _grads(z, img_lo, img_hi, denom, S=2):
    from oracle import ort_oracle as O
    cfg, full = z["cfg"], z["w"]
    W = {k: v.clone().requires_grad_(True) for k, v in full.items()
         if v.is_floating_point() and not k.endswith(".pe") and not k.endswith("_pruning_mask")}
    Sg = {k: full[k + "_pruning_mask"].clone().requires_grad_(True) for k in z["u"]}
    eff = dict(W)
    for k in z["u"]:
        p = torch.sigmoid(Sg[k])
        m = (z["u"][k] < p).float()
        eff[k] = (p + (m - p).detach()) * W[k]
    eff["model.tgt_embed.1.pe"] = full["model.tgt_embed.1.pe"]
    rows = slice(img_lo * S, img_hi * S)
    lp = O.forward_tf(eff, cfg, z["att_feats"][img_lo:img_hi], z["boxes"][img_lo:img_hi], z["seqs"][rows], None)
    tgt, msk = z["seqs"][rows, 1:], z["masks"][rows, 1:]
    loss = -(lp.gather(2, tgt.unsqueeze(2)).squeeze(2) * msk).sum() / denom
    loss.backward()
    names = sorted(W)
    flat_w = torch.cat([W[k].grad.reshape(-1) for k in names])
    flat_s = torch.cat([Sg[k].grad.reshape(-1) for k in sorted(Sg)])
    return flat_w, flat_s


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from sparse_caption_b200 import distributed as D
    from tests import golden_io
    z = golden_io.load("ort_prune_tiny")
    B = z["att_feats"].shape[0]
    lo, hi = D.shard_range(B, rank, world)
    T = z["seqs"].shape[1] - 1
    denom = D.global_token_count(z["masks"][lo * 2: hi * 2], T)
    gw, gs = _grads(z, lo, hi, denom)
    # weights: bucketed + asynchronous (what OrtTrainer does per backward phase); logits: one blocking call
    ar_async = D.make_all_reduce(async_op=True)
    cut = gw.numel() // 3
    handles = [ar_async(gw[:cut]), ar_async(gw[cut:])]
    for h in handles:
        h.wait()
    ar = D.make_all_reduce()
    assert ar(gs) is None
    if rank == 0:
        torch.save({"gw": gw, "gs": gs, "denom": denom}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradients_equal_single_process(tmp_path):
    sys.path.insert(0, ROOT)
    from sparse_caption_b200 import distributed as D
    assert [D.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert D.make_all_reduce() is None  # not initialised -> single process
    out = str(tmp_path / "ddp.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    from tests import golden_io
    z = golden_io.load("ort_prune_tiny")
    T = z["seqs"].shape[1] - 1
    denom = z["masks"][:, 1: T + 1].sum()
    assert float(got["denom"]) == float(denom)
    gw, gs = _grads(z, 0, z["att_feats"].shape[0], denom)
    torch.testing.assert_close(got["gw"], gw, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(got["gs"], gs, rtol=1e-4, atol=1e-7)


def _sharded_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sparse_caption_b200 import distributed as D
    g = torch.Generator().manual_seed(100 + rank)
    n = 1003  # not a multiple of the world size: exercises the redundantly updated tail
    param = torch.arange(n, dtype=torch.float32) / 7.0          # identical on every rank
    grad = torch.randn(n, generator=g)                          # rank-local gradient
    ex = D.ShardedExchange()
    calls = []

    def update(lo, hi):  # stand-in optimizer: SGD with lr 0.5 on the rank-summed gradient
        calls.append((lo, hi))
        param[lo:hi] -= 0.5 * grad[lo:hi]

    for a, b in ((0, 400), (400, 401), (401, n)):  # three buckets, one shorter than the world size
        ex.bucket(grad, param, a, b, update)
    ex.finish()
    torch.save({"param": param, "grad_local_seed": 100 + rank, "calls": calls}, f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_exchange_equals_allreduce_update(tmp_path):
    """reduce-scatter -> owner update -> all-gather (distributed.ShardedExchange) leaves every rank with the parameters an
    all-reduce + full update would produce; each element is updated by exactly one rank, tails by all."""
    out = str(tmp_path / "zero1.pt")
    port = 31500 + os.getpid() % 2000
    world = 3
    mp.spawn(_sharded_worker, args=(world, port, out), nprocs=world, join=True)
    n = 1003
    total = sum(torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in range(world))
    want = torch.arange(n, dtype=torch.float32) / 7.0 - 0.5 * total
    got = [torch.load(f"{out}.{r}") for r in range(world)]
    for r in range(world):
        torch.testing.assert_close(got[r]["param"], want, rtol=1e-6, atol=1e-6)
    # ownership: the shards of bucket (0, 400) partition [0, 399) and the tail [399, 400) is updated by everyone
    owned = sorted(c for r in range(world) for c in got[r]["calls"] if c[1] <= 399)
    assert owned == [(0, 133), (133, 266), (266, 399)]
    assert all((399, 400) in got[r]["calls"] and (400, 401) in got[r]["calls"] for r in range(world))


def _dwm_worker(rank, world, port, out):
    """Rank-local dWm = dL/d(W.m) (gradient with respect to the MASKED weight), summed over ranks; rank 0 then applies the
    straight-through epilogue once - what OrtTrainer(fused_st=True) + sc_adam_clip_st do on the device."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import ort_oracle as O
    from sparse_caption_b200 import distributed as D
    from tests import golden_io
    z = golden_io.load("ort_prune_tiny")
    cfg, full = z["cfg"], z["w"]
    B = z["att_feats"].shape[0]
    lo, hi = D.shard_range(B, rank, world)
    T = z["seqs"].shape[1] - 1
    denom = D.global_token_count(z["masks"][lo * 2: hi * 2], T)
    keys = sorted(z["u"])
    eff = {k: v.clone() for k, v in full.items() if not k.endswith("_pruning_mask")}
    for k in keys:  # same Bernoulli sample on every rank (same uniforms / same Philox seed)
        m = (z["u"][k] < torch.sigmoid(full[k + "_pruning_mask"])).float()
        eff[k] = (m * full[k]).requires_grad_(True)
    rows = slice(lo * 2, hi * 2)
    lp = O.forward_tf(eff, cfg, z["att_feats"][lo:hi], z["boxes"][lo:hi], z["seqs"][rows], None)
    loss = -(lp.gather(2, z["seqs"][rows, 1:].unsqueeze(2)).squeeze(2) * z["masks"][rows, 1:]).sum() / denom
    loss.backward()
    dwm = torch.cat([eff[k].grad.reshape(-1) for k in keys])
    ar = D.make_all_reduce()
    ar(dwm)                                                       # ONE buffer on the wire (vs dW and dS)
    if rank == 0:
        torch.save({"dwm": dwm}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_dwm_exchange_equals_dw_and_ds_exchange(tmp_path):
    """All-reducing dWm once and forming dW = dWm (.) m, dS = dWm (.) W (.) sigmoid'(S) afterwards equals all-reducing dW and
    dS (both are linear in dWm with rank-independent factors): half the bytes, same update."""
    out = str(tmp_path / "dwm.pt")
    port = 33500 + os.getpid() % 2000
    mp.spawn(_dwm_worker, args=(2, port, out), nprocs=2, join=True)
    from tests import golden_io
    z = golden_io.load("ort_prune_tiny")
    keys = sorted(z["u"])
    dwm = torch.load(out)["dwm"]
    T = z["seqs"].shape[1] - 1
    gw_ref, gs_ref = _grads(z, 0, z["att_feats"].shape[0], z["masks"][:, 1: T + 1].sum())
    names = sorted(k for k, v in z["w"].items() if v.is_floating_point() and not k.endswith(".pe") and not k.endswith("_pruning_mask"))
    ref_w = dict(zip(names, torch.split(gw_ref, [z["w"][k].numel() for k in names])))
    ref_s = dict(zip(keys, torch.split(gs_ref, [z["w"][k].numel() for k in keys])))
    off = 0
    for k in keys:
        n = z["w"][k].numel()
        g = dwm[off: off + n].view(z["w"][k].shape)
        off += n
        S, W = z["w"][k + "_pruning_mask"], z["w"][k]
        sig = torch.sigmoid(S)
        m = (z["u"][k] < sig).float()
        torch.testing.assert_close(g * m, ref_w[k].view(W.shape), rtol=1e-4, atol=1e-7)
        torch.testing.assert_close(g * W * sig * (1 - sig), ref_s[k].view(W.shape), rtol=1e-4, atol=1e-7)
