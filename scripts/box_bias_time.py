"""sc_box_bias_all (fp32 FFMA) vs sc_box_bias_all_tc (TF32 mma) and the encoder attention at the coalesced encoder batch."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from bench import _time_graph
from sparse_caption_b200 import kernels as K, lib
lib.load()
dev = torch.device("cuda", 0)
B, N, h, d = int(os.environ.get("SC_IMAGES", "2560")), 36, 8, 512
boxes = torch.rand(B, N, 4, device=dev); boxes[..., 2:] += boxes[..., :2]
wgw = torch.randn(6 * h, 64, device=dev) * 0.1; wgb = torch.rand(6 * h, device=dev)
bias = torch.empty(6, B, h, N, N, device=dev)
for tc in (False, True):
    us = _time_graph(lambda i: K.box_bias_all(boxes, wgw, wgb, bias, B=B, N=N, layers=6, h=h, tensor_cores=tc), dev, reps=10)
    print(f"box_bias_all tensor_cores={tc}: {us:8.1f} us  ({bias.numel() * 4 / us / 1e3:.0f} GB/s written)")
qkv = torch.randn(B * N, 3 * d, device=dev).bfloat16(); out = torch.empty(B * N, d, device=dev, dtype=torch.bfloat16)
us = _time_graph(lambda i: K.bias_attention(qkv[:, 0:], qkv[:, d:], qkv[:, 2 * d:], bias[i % 6], None, out, B=B, N=N, h=h, dk=64, ldq=3 * d,
                                            ldk=3 * d, ldv=3 * d, ldo=d), dev, reps=12)
nb = B * N * 4 * d * 2 + B * h * N * N * 4
print(f"enc attention: {us:8.1f} us  {nb / us / 1e3:.0f} GB/s (algorithmic {nb / 1e6:.0f} MB)")
