#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do echo "default: $(SC_WALL_ONLY=1 python scripts/profile_train.py 2>&1 | tail -1)"; done
echo "ring=2: $(SC_WALL_ONLY=1 SC_WGRAD_RING=2 python scripts/profile_train.py 2>&1 | tail -1)"
echo "ring=6: $(SC_WALL_ONLY=1 SC_WGRAD_RING=6 python scripts/profile_train.py 2>&1 | tail -1)"
SC_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python scripts/profile_train.py > gpurun_out/ncu_train.log 2>&1
python scripts/ncu_agg.py gpurun_out/launches_train.csv 40 | tee gpurun_out/launches_train_summary.txt
