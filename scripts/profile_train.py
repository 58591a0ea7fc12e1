"""Per-kernel / per-shape CUDA-event timing of one SMP training step (diagnostic; not a bench number)."""
import os, sys, collections, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sparse_caption_b200 import lib, synthetic
from sparse_caption_b200.engine import ModelCfg
from sparse_caption_b200.trainer import OrtTrainer
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 50
dev = torch.device("cuda")
cfg = ModelCfg(dict(bench.CFG, max_seq_length=17))
sd = synthetic.random_state_dict(cfg, seed=1234, sparsity=0.0, device=dev)
GRAPH = os.environ.get("SC_TRAIN_GRAPH", "1") == "1"
tr = OrtTrainer(sd, cfg, mask_type="supermask", precision="bf16", device=dev, seed=8888, use_graph=GRAPH)
if "SC_WGRAD_RING" in os.environ:
    tr.wgrad_ring = int(os.environ["SC_WGRAD_RING"])
if "SC_PDL_MASK" in os.environ:
    tr.pdl_mask = int(os.environ["SC_PDL_MASK"])
if os.environ.get("SC_SKIP_WGRAD") == "1":
    tr._diag_skip_wgrad = True   # (wrong gradients: timing of the main chain alone)
S, T = 5, 17
g = torch.Generator().manual_seed(8888)
att, boxes = synthetic.synthetic_inputs(B, 36, 2048, seed=8888, pin=True)
R = B * S
seqs = torch.zeros(R, T + 1, dtype=torch.long); masks = torch.zeros(R, T + 1)
lens = torch.randint(6, T - 1, (R,), generator=g)
for r in range(R):
    n = int(lens[r]); seqs[r, 0] = 2; seqs[r, 1:1 + n] = torch.randint(4, 10000, (n,), generator=g); seqs[r, 1 + n] = 3; masks[r, :n + 2] = 1
opt = dict(lr=3e-4, sparsity_target=0.95, sparsity_weight=30.0, current_step=100, max_step=1000)
for _ in range(3):
    tr.train_step(att, boxes, seqs, masks, seq_per_img=S, **opt)
torch.cuda.synchronize()
if os.environ.get("SC_NCU_RANGE") == "1":
    # ncu --profile-from-start off: exactly one eager step inside the profiler range
    tr.use_graph = False
    tr.train_step(att, boxes, seqs, masks, seq_per_img=S, **opt); torch.cuda.synchronize()
    torch.cuda.profiler.start()
    tr.train_step(att, boxes, seqs, masks, seq_per_img=S, **opt); torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
t0 = time.perf_counter()
for _ in range(5):
    tr.train_step(att, boxes, seqs, masks, seq_per_img=S, **opt)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue ms/step {(t1-t0)/5*1e3:.2f}  wall ms/step {(t2-t0)/5*1e3:.2f}")
if os.environ.get("SC_WALL_ONLY") == "1":
    kw = {}
    if os.environ.get("SC_FAKE_AR") == "1":   # phase-split graphs + a no-op exchange: cost of splitting the step
        kw["all_reduce"] = lambda t: None
    for _ in range(3):
        tr.train_step(att, boxes, seqs, masks, seq_per_img=S, **kw, **opt)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(40):
        tr.train_step(att, boxes, seqs, masks, seq_per_img=S, **kw, **opt)
    torch.cuda.synchronize()
    print(f"wall ms/step over 40 steps {(time.perf_counter()-t0)/40*1e3:.3f}")
    sys.exit(0)
tr.use_graph = False  # per-kernel event timing needs the eager launch sequence
agg = collections.defaultdict(lambda: [0.0, 0])
for rep in range(3):
    lib.profile = []
    tr.train_step(att, boxes, seqs, masks, seq_per_img=S, **opt); torch.cuda.synchronize()
    prof, lib.profile = lib.profile, None
    for name, meta, a, b in prof:
        key = (name,) + (tuple(meta[:4]) if meta else ())
        agg[key][0] += a.elapsed_time(b) / 3; agg[key][1] += 1
rows = sorted(agg.items(), key=lambda kv: -kv[1][0])
tot = sum(v[0] for v in agg.values())
print(f"total kernel ms/step {tot:.3f}")
for k, (ms, n) in rows:
    n //= 3
    extra = ""
    if len(k) > 1 and str(k[1]).startswith("gemm"):
        M, N, K = k[2], k[3], k[4]
        extra = f" {2*M*N*K*n/ms/1e9:8.1f} TF/s"
    print(f"{ms:8.3f} ms  n={n:4d}  avg={1e3*ms/max(n,1):8.1f} us  {k}{extra}")
