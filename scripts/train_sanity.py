"""Full-size SMP training sanity: the graph-replayed step (all fused paths on) overfits one fixed batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from sparse_caption_b200 import synthetic
from sparse_caption_b200.engine import ModelCfg
from sparse_caption_b200.trainer import OrtTrainer
dev = torch.device("cuda")
cfg = ModelCfg(dict(bench.CFG, max_seq_length=17))
sd = synthetic.random_state_dict(cfg, seed=1234, sparsity=0.0, device=dev)
B, S, T = 50, 5, 17
g = torch.Generator().manual_seed(1)
att, boxes = synthetic.synthetic_inputs(B, 36, 2048, seed=2, pin=True)
R = B * S
seqs = torch.zeros(R, T + 1, dtype=torch.long); masks = torch.zeros(R, T + 1)
for r in range(R):
    n = int(torch.randint(6, T - 1, (1,), generator=g)); seqs[r, 0] = 2
    seqs[r, 1:1 + n] = torch.randint(4, 10000, (n,), generator=g); seqs[r, 1 + n] = 3; masks[r, :n + 2] = 1
for use_graph in (True, False):
    tr = OrtTrainer(sd, cfg, mask_type="supermask", precision="bf16", device=dev, seed=5, use_graph=use_graph)
    losses = []
    for i in range(60):
        losses.append(float(tr.train_step(att, boxes, seqs, masks, seq_per_img=S, lr=5e-4, sparsity_target=0.8, sparsity_weight=30.0,
                                          current_step=i, max_step=60)))
    kept = float((tr.flat_s > 0).float().mean())
    print(f"graph={use_graph}: loss step 1 {losses[0]:.3f}, 10 {losses[9]:.3f}, 30 {losses[29]:.3f}, 60 {losses[59]:.3f}; "
          f"mask logits > 0: {kept * 100:.1f}%  finite={all(l == l for l in losses)}")
    assert losses[59] < losses[0] - 1.0, "the loss did not go down"
