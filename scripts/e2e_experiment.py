"""H2D bandwidth, host fp32->bf16 conversion speed, and the end-to-end step with bf16 host features (upper bound)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sparse_caption_b200 import synthetic
from sparse_caption_b200.engine import ModelCfg, OrtEngine
dev = torch.device("cuda")
B = 512
att, boxes = synthetic.synthetic_inputs(B, 36, 2048, seed=1, pin=True)
d = torch.empty(B, 36, 2048, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
d.copy_(att, non_blocking=True); torch.cuda.synchronize()
e0.record()
for _ in range(5): d.copy_(att, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print(f"H2D pinned fp32 151 MB: {att.numel()*4*5/e0.elapsed_time(e1)/1e6:.1f} GB/s")
stage = torch.empty(B, 36, 2048, dtype=torch.bfloat16).pin_memory()
for nt in (4, 8, 16, os.cpu_count()):
    torch.set_num_threads(nt)
    stage.copy_(att)
    t0 = time.perf_counter()
    for _ in range(5): stage.copy_(att)
    print(f"host fp32->bf16 with {nt} threads: {(time.perf_counter()-t0)/5*1e3:.2f} ms per batch")
cfg = ModelCfg(bench.CFG)
sd = synthetic.random_state_dict(cfg, seed=1234, sparsity=0.95, device=dev)
for tag, feats in (("fp32 host features", att), ("bf16 host features", stage)):
    eng = OrtEngine(sd, cfg, precision="bf16", device=dev, dec_tiles={"o": 3256, "co": 3256, "cq": 3128, "ff2": 3256})
    S = 8
    outs = [(torch.empty(B, 3, 16, dtype=torch.int32).pin_memory(), torch.empty(B, 3, 16).pin_memory()) for _ in range(S)]
    for i in range(S): eng.submit(feats, boxes, None, {"beam_size": 3}, slot=i + 1, out=outs[i])
    eng.wait(); torch.cuda.synchronize()
    e0.record()
    for i in range(16): eng.submit(feats, boxes, None, {"beam_size": 3}, slot=i % S + 1, out=outs[i % S])
    eng.wait(); e1.record(); torch.cuda.synchronize()
    print(f"e2e {tag}: {e0.elapsed_time(e1)/16:.3f} ms/step")
    del eng
