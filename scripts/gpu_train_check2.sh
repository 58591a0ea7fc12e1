#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_train_kernels_gpu.py tests/test_trainer_gpu.py tests/test_dropin_gpu.py -q -m gpu > gpurun_out/t_train.log 2>&1; echo "tests exit=$? $(tail -n 1 gpurun_out/t_train.log)"
grep -E "^E  |Error|FAILED" gpurun_out/t_train.log | head -30
timeout -s KILL 300 python scripts/profile_train.py > gpurun_out/profile_train.txt 2>&1; head -3 gpurun_out/profile_train.txt
SC_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python scripts/profile_train.py > gpurun_out/ncu_train.log 2>&1
python scripts/ncu_agg.py gpurun_out/launches_train.csv 45 | tee gpurun_out/launches_train_summary.txt
