#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "linear" > gpurun_out/t_linear.log 2>&1; echo "linear tests exit=$? $(tail -n 1 gpurun_out/t_linear.log)"
timeout -s KILL 600 python -m pytest tests/test_train_kernels_gpu.py -q -m gpu -x -k "masked_linear_backward" > gpurun_out/t_wgrad.log 2>&1; echo "wgrad tests exit=$? $(tail -n 1 gpurun_out/t_wgrad.log)"
grep -E "^E  |Error|FAILED" gpurun_out/t_linear.log gpurun_out/t_wgrad.log | head -30
timeout -s KILL 300 python scripts/ln_sweep.py 2>&1 | tee gpurun_out/ln_sweep.txt
timeout -s KILL 300 python scripts/gemm_sweep.py 1536,512,512 1536,1536,512 1536,2048,512 1536,512,2048 1536,10000,512 18432,512,2048 18432,1536,512 18432,2048,512 18432,512,512 4250,512,512 4250,2048,512 4250,10000,512 2>&1 | tee gpurun_out/gemm_sweep3.txt
