#!/bin/bash
# compute-sanitizer memcheck over the kernels added this round (small shapes only)
mkdir -p gpurun_out
timeout -s KILL 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_train_kernels_gpu.py tests/test_kernels_gpu.py -q -m gpu -x \
  -k "tensor_path or hmask or batched or sell_spmm or topk_records or partials or sample_step or layernorm_bwd or prep_grad or box_bias" > gpurun_out/sanitize.log 2>&1
echo "memcheck exit=$?"; grep -E "ERROR SUMMARY|Invalid|out of bounds|passed|failed" gpurun_out/sanitize.log | head -20
