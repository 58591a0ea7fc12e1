"""Fused ingest kernel (pinned host fp32 -> device bf16 over PCIe) vs cudaMemcpyAsync + cast: bandwidth and correctness."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_caption_b200.kernels as K
from sparse_caption_b200 import synthetic
att, boxes = synthetic.synthetic_inputs(512, 36, 2048, seed=1, pin=True)
dev = torch.device("cuda")
out = torch.empty(512 * 36 * 2048, device=dev, dtype=torch.bfloat16)
ref = att.to(dev).bfloat16().view(-1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for ctas in (16, 32, 64, 148, 296):
    out.zero_(); K.ingest_f32_bf16(att, out, ctas=ctas); torch.cuda.synchronize()
    ok = torch.equal(out, ref)
    e0.record()
    for _ in range(5): K.ingest_f32_bf16(att, out, ctas=ctas)
    e1.record(); torch.cuda.synchronize()
    print(f"ingest kernel, {ctas:3d} CTAs: {att.numel() * 4 * 5 / e0.elapsed_time(e1) / 1e6:.1f} GB/s over PCIe, exact={ok}")
