"""In-graph per-kernel durations and idle gaps of the inference step via torch.profiler (CUPTI), diagnostic only."""
import os, sys, collections, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from sparse_caption_b200 import lib, synthetic
from sparse_caption_b200.engine import ModelCfg, OrtEngine
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfg = ModelCfg(bench.CFG)
dev = torch.device("cuda")
sd = synthetic.random_state_dict(cfg, seed=1234, sparsity=bench.SPARSITY, device=dev)
eng = OrtEngine(sd, cfg, precision="bf16", sparse_backend="dense", device=dev)
att, boxes = synthetic.synthetic_inputs(B, 36, 2048, seed=1, pin=True)
opt = {"beam_size": 3}
enc = eng.encode(att, boxes); eng.decode(enc, opt); torch.cuda.synchronize()
for _ in range(2):
    eng.run_encoder(enc); eng.decode(enc, opt)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    eng.run_encoder(enc); eng.decode(enc, opt)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
agg = collections.defaultdict(lambda: [0.0, 0])
busy = 0.0
gaps = 0.0
prev_end = None
for e in evs:
    d = e.time_range.end - e.time_range.start
    name = e.name[:70]
    agg[name][0] += d; agg[name][1] += 1
    busy += d
    if prev_end is not None and e.time_range.start > prev_end:
        gaps += e.time_range.start - prev_end
    prev_end = max(prev_end or 0, e.time_range.end)
span = evs[-1].time_range.end - evs[0].time_range.start
print(f"kernels {len(evs)} span {span/1e3:.3f} ms busy {busy/1e3:.3f} ms gaps {gaps/1e3:.3f} ms")
for k, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{us/1e3:8.3f} ms  n={n:4d}  avg={us/n:8.2f} us  {k}")
