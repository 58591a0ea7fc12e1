#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_dropin_gpu.py tests/test_trainer_gpu.py -q -m gpu 2>&1 | grep -E "^E  |Error|passed|failed|FAILED" | head -30
