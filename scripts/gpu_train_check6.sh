#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_train_kernels_gpu.py tests/test_trainer_gpu.py -q -m gpu > gpurun_out/t_train.log 2>&1; echo "tests exit=$? $(tail -n 1 gpurun_out/t_train.log)"
grep -E "^E  |Error|FAILED" gpurun_out/t_train.log | head -30
echo "both fusions: $(SC_WALL_ONLY=1 python scripts/profile_train.py 2>&1 | tail -1)"
echo "no hmask: $(SC_WALL_ONLY=1 SC_NO_HMASK=1 python scripts/profile_train.py 2>&1 | tail -1)"
echo "no attn16: $(SC_WALL_ONLY=1 SC_NO_ATTN16=1 python scripts/profile_train.py 2>&1 | tail -1)"
echo "neither: $(SC_WALL_ONLY=1 SC_NO_HMASK=1 SC_NO_ATTN16=1 python scripts/profile_train.py 2>&1 | tail -1)"
echo "ring=1: $(SC_WALL_ONLY=1 SC_WGRAD_RING=1 python scripts/profile_train.py 2>&1 | tail -1)"
