#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "linear" > gpurun_out/t_linear.log 2>&1; echo "linear tests exit=$? $(tail -n 1 gpurun_out/t_linear.log)"
timeout -s KILL 600 python -m pytest tests/test_train_kernels_gpu.py -q -m gpu -x -k "masked_linear_backward" > gpurun_out/t_wgrad.log 2>&1; echo "wgrad tests exit=$? $(tail -n 1 gpurun_out/t_wgrad.log)"
timeout -s KILL 300 python scripts/gemm_sweep.py > gpurun_out/gemm_sweep2.txt 2>&1; echo "sweep exit=$?"
cat gpurun_out/gemm_sweep2.txt
grep -E "^E  |Error|FAILED" gpurun_out/t_linear.log gpurun_out/t_wgrad.log | head -30
