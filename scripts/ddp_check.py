"""torchrun --nproc-per-node 2 scripts/ddp_check.py : the sharded exchange (reduce-scatter -> Adam on the owned shard ->
all-gather) against the all-reduce + full optimizer path: same parameters after several SMP steps on every rank."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from sparse_caption_b200 import distributed as D, synthetic
from sparse_caption_b200.engine import ModelCfg
from sparse_caption_b200.trainer import OrtTrainer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = ModelCfg(dict(d_model=512, dim_feedforward=1024, num_layers=2, num_heads=8, max_seq_length=9, att_feat_size=128, vocab_size=500))
sd = synthetic.random_state_dict(cfg, seed=3, sparsity=0.0, device=dev)
B, S, T = 8, 2, 9
g = torch.Generator().manual_seed(50 + rank)
att, boxes = synthetic.synthetic_inputs(B, 36, 128, seed=60 + rank, pin=True)
seqs = torch.zeros(B * S, T + 1, dtype=torch.long); masks = torch.zeros(B * S, T + 1)
for r in range(B * S):
    n = int(torch.randint(3, T - 1, (1,), generator=g)); seqs[r, 0] = 2
    seqs[r, 1:1 + n] = torch.randint(4, 500, (n,), generator=g); seqs[r, 1 + n] = 3; masks[r, :n + 2] = 1
opt = dict(lr=1e-3, sparsity_target=0.9, sparsity_weight=5.0, current_step=3, max_step=10)
gtok = D.global_token_count(masks.to(dev), T)
res = {}
# one step: bias / LayerNorm gradients and the loss are atomic sums (order-dependent last bits) and Adam's first update is
# lr * g / |g|, so a handful of near-zero gradients may flip sign between ANY two runs; everything else must agree exactly
NSTEPS = 1
for mode in ("allreduce", "sharded"):
    for use_graph in (False, True):
        tr = OrtTrainer(sd, cfg, mask_type="supermask", precision="bf16", device=dev, seed=77, use_graph=use_graph)
        kw = dict(all_reduce=D.make_all_reduce(async_op=True)) if mode == "allreduce" else dict(exchange=D.ShardedExchange(device=dev))
        losses = [float(tr.train_step(att, boxes, seqs, masks, seq_per_img=S, global_tokens=gtok, **kw, **opt)) for _ in range(NSTEPS)]
        torch.cuda.synchronize()
        res[(mode, use_graph)] = (losses, tr.flat_w.clone(), tr.flat_s.clone())
ok = True
for use_graph in (False, True):
    la, wa, sa = res[("allreduce", use_graph)]
    ls, ws_, ss = res[("sharded", use_graph)]
    dw = float((wa - ws_).abs().max()); ds = float((sa - ss).abs().max())
    # every rank must hold the same parameters as rank 0
    w0 = ws_.clone(); dist.broadcast(w0, 0); s0 = ss.clone(); dist.broadcast(s0, 0)
    same = torch.equal(w0, ws_) and torch.equal(s0, ss)
    print(f"rank {rank} graph={use_graph}: losses allreduce {la} sharded {ls}  max|dW| {dw:.3e} max|dS| {ds:.3e} ranks identical {same}", flush=True)
    fw = float(((wa - ws_).abs() > 1e-6).float().mean()); fs = float(((sa - ss).abs() > 1e-3).float().mean())
    print(f"rank {rank} graph={use_graph}: fraction of weights differing {fw:.2e}, of logits {fs:.2e}", flush=True)
    ok = ok and fw < 1e-3 and fs < 1e-3 and same and all(abs(a - b) < 1e-4 for a, b in zip(la, ls))
print(f"rank {rank}: {'DDP CHECK OK' if ok else 'DDP CHECK FAILED'}", flush=True)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
