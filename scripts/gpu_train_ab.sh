#!/bin/bash
# training-step A/B on one box: decoder-side batched mask pass underneath the encoder forward (default) vs in front of it
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_trainer_gpu.py tests/test_parity_baseline_gpu.py -q -x -k "train or grad or smp" 2>&1 | tail -3
for v in 0 1 0 1; do
  echo "premask_overlap=$v: $(SC_PREMASK_OVERLAP=$v SC_WALL_ONLY=1 python scripts/profile_train.py 2>&1 | tail -1)"
done
