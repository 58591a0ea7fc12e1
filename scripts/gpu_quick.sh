#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_trainer_gpu.py -q -m gpu -s -k "all_gradients" 2>&1 | grep -E "^E  |Error|trainer.py:[0-9]+|passed|failed|FAILED|bf16|    model|    att" | head -40
