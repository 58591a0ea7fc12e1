"""torchrun --nproc-per-node N scripts/ddp_time.py : wall ms/step of the data-parallel SMP training step (bench shapes)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from sparse_caption_b200 import distributed as D, synthetic
from sparse_caption_b200.engine import ModelCfg
from sparse_caption_b200.trainer import OrtTrainer
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = ModelCfg(dict(bench.CFG, max_seq_length=17))
sd = synthetic.random_state_dict(cfg, seed=1234, sparsity=0.0, device=dev)
tr = OrtTrainer(sd, cfg, mask_type="supermask", precision="bf16", device=dev, seed=8888, use_graph=True)
B, S, T = 50, 5, 17
g = torch.Generator().manual_seed(8888 + rank)
att, boxes = synthetic.synthetic_inputs(B, 36, 2048, seed=8888 + rank, pin=True)
R = B * S
seqs = torch.zeros(R, T + 1, dtype=torch.long); masks = torch.zeros(R, T + 1)
lens = torch.randint(6, T - 1, (R,), generator=g)
for r in range(R):
    n = int(lens[r]); seqs[r, 0] = 2; seqs[r, 1:1 + n] = torch.randint(4, 10000, (n,), generator=g); seqs[r, 1 + n] = 3; masks[r, :n + 2] = 1
seqs, masks = seqs.pin_memory(), masks.pin_memory()
opt = dict(lr=3e-4, sparsity_target=0.95, sparsity_weight=30.0, current_step=100, max_step=1000)
gtok = D.global_token_count(masks.to(dev), T)
mode = os.environ.get("SC_EXCHANGE", "allreduce")
kw = dict(exchange=D.ShardedExchange(device=dev)) if mode == "sharded" else dict(all_reduce=D.make_all_reduce(async_op=True))
if mode == "none": kw = dict(all_reduce=lambda t: None)
for _ in range(4): tr.train_step(att, boxes, seqs, masks, seq_per_img=S, global_tokens=gtok, **kw, **opt)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(40): tr.train_step(att, boxes, seqs, masks, seq_per_img=S, global_tokens=gtok, **kw, **opt)
torch.cuda.synchronize(); dist.barrier()
if rank == 0: print(f"N={world} exchange={mode} NCCL_MAX_NCHANNELS={os.environ.get('NCCL_MAX_NCHANNELS')} NCCL_ALGO={os.environ.get('NCCL_ALGO')}: {(time.perf_counter()-t0)/40*1e3:.3f} ms/step", flush=True)
dist.destroy_process_group()
