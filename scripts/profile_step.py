"""Per-kernel / per-shape CUDA-event timing of one un-graphed inference step (diagnostic; not a bench number)."""
import os, sys, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sparse_caption_b200 import lib, synthetic
from sparse_caption_b200.engine import ModelCfg, OrtEngine
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
backend = sys.argv[2] if len(sys.argv) > 2 else "dense"
cfg = ModelCfg(bench.CFG)
dev = torch.device("cuda")
sd = synthetic.random_state_dict(cfg, seed=1234, sparsity=bench.SPARSITY, device=dev)
eng = OrtEngine(sd, cfg, precision="bf16", sparse_backend=backend, device=dev, use_graphs=False)
att, boxes = synthetic.synthetic_inputs(B, 36, 2048, seed=1, pin=True)
opt = {"beam_size": 3}
enc = eng.encode(att, boxes); eng.decode(enc, opt); torch.cuda.synchronize()
if os.environ.get("SC_NCU_RANGE") == "1":
    # ncu --profile-from-start off: exactly one eager step inside the profiler range
    torch.cuda.profiler.start()
    eng.run_encoder(enc); eng.decode(enc, opt); torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
agg = collections.defaultdict(lambda: [0.0, 0])
for rep in range(3):
    lib.profile = []
    eng.run_encoder(enc); eng.decode(enc, opt); torch.cuda.synchronize()
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
    if len(k) > 1 and k[1].startswith("gemm"):
        M, N, K = k[2], k[3], k[4]
        extra = f" {2*M*N*K*n/ms/1e9:8.1f} TF/s"
    print(f"{ms:8.3f} ms  n={n:4d}  avg={1e3*ms/n:8.1f} us  {k}{extra}")
