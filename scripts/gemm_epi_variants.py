"""Where the N = d_model GEMM's time goes: 7680 x 512 x {512, 2048} with each combination of output dtype / residual, alone in a graph."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from bench import _time_graph
from sparse_caption_b200 import kernels as K, lib
lib.load()
dev = torch.device("cuda", 0)
R = int(os.environ.get("SC_ROWS", "7680"))
for Kd in (512, 2048):
    x = torch.randn(R, Kd, device=dev).bfloat16(); w = torch.randn(512, Kd, device=dev).bfloat16(); b = torch.randn(512, device=dev)
    x32 = torch.randn(R, 512, device=dev); y32 = torch.empty(R, 512, device=dev); y16 = torch.empty(R, 512, device=dev, dtype=torch.bfloat16)
    for hint in (3256, 20003256, 3128, 5128):
        for name, res, out in (("bf16 out        ", None, y16), ("fp32 out        ", None, y32), ("bf16 out + res  ", x32, y16),
                               ("fp32 out + res  ", x32, y32), ("fp32 in place   ", x32, x32)):
            try:
                us = _time_graph(lambda i: K.linear(x, w, b, residual=res, out=out, tile_n=hint), dev)
                print(f"K={Kd:5d} hint={hint:9d} {name} {us:7.2f} us  {2.0*R*512*Kd/us/1e6:6.0f} TF/s", flush=True)
            except Exception as ex:
                print("#", hint, name, ex)
