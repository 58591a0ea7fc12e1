#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_trainer_gpu.py -q -m gpu > gpurun_out/t_s.log 2>&1; echo "tests exit=$? $(tail -n 1 gpurun_out/t_s.log)"
grep -E "^E  |Error|FAILED" gpurun_out/t_s.log | head -30
