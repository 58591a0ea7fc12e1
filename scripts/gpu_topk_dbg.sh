#!/bin/bash
for d in 0 1 2; do SC_GEMM_DBG=$d python scripts/dec_kernels.py --hints 20003256 2>&1 | grep -E "topk" | sed "s/^/dbg=$d /"; done
python - <<'PY'
import torch, sys, os
sys.path.insert(0, '.')
from bench import _time_graph
from sparse_caption_b200 import kernels as K, lib
lib.load()
dev = torch.device('cuda', 0)
R = 7680
x = torch.randn(R, 512, device=dev).bfloat16(); w = torch.randn(10000, 512, device=dev).bfloat16(); b = torch.randn(10000, device=dev)
y16 = torch.empty(R, 10000, device=dev, dtype=torch.bfloat16)
for hint in (0, 3256, 20003256):
    us = _time_graph(lambda i: K.linear(x, w, b, out=y16, tile_n=hint), dev)
    print(f"plain generator GEMM bf16 out hint={hint}: {us:.1f} us  {2.0*R*10000*512/us/1e6:.0f} TF/s")
PY
