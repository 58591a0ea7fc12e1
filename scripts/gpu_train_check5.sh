#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_train_kernels_gpu.py tests/test_trainer_gpu.py tests/test_dropin_gpu.py -q -m gpu > gpurun_out/t_train.log 2>&1; echo "tests exit=$? $(tail -n 1 gpurun_out/t_train.log)"
grep -E "^E  |Error|FAILED" gpurun_out/t_train.log | head -30
for r in 4 6 8; do echo "ring=$r: $(SC_WALL_ONLY=1 SC_WGRAD_RING=$r python scripts/profile_train.py 2>&1 | tail -1)"; done
echo "ring=1: $(SC_WALL_ONLY=1 SC_WGRAD_RING=1 python scripts/profile_train.py 2>&1 | tail -1)"
