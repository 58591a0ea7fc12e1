#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_train_kernels_gpu.py tests/test_trainer_gpu.py tests/test_dropin_gpu.py -q -m gpu -x > gpurun_out/t_train.log 2>&1; echo "train tests exit=$? $(tail -n 1 gpurun_out/t_train.log)"
grep -E "^E  |Error|FAILED" gpurun_out/t_train.log | head
timeout -s KILL 300 python scripts/profile_train.py > gpurun_out/profile_train.txt 2>&1; head -45 gpurun_out/profile_train.txt
