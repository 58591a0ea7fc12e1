#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python -m pytest tests/test_train_kernels_gpu.py tests/test_kernels_gpu.py -q -m gpu -x \
  -k "tensor_path or batched or sell_spmm or layernorm_bwd or prep_grad or box_bias or partials" > gpurun_out/racecheck.log 2>&1
echo "racecheck exit=$?"; grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/racecheck.log | head -20
timeout -s KILL 600 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_train_kernels_gpu.py -q -m gpu -x -k "tensor_path or hmask or layernorm_bwd" > gpurun_out/synccheck.log 2>&1
echo "synccheck exit=$?"; grep -E "ERROR SUMMARY|Barrier|divergent|passed|failed" gpurun_out/synccheck.log | head -10
