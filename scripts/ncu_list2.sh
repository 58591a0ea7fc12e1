#!/bin/bash
# launch lists (cold-cache, serialised: compare SHARES) of one eager training step and one un-graphed inference step
mkdir -p gpurun_out
SC_TRAIN_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 720 --csv --log-file gpurun_out/launches_train.csv python scripts/profile_train.py > gpurun_out/ncu_train.log 2>&1
echo "ncu train exit=$?"; wc -l gpurun_out/launches_train.csv
ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 1000 --csv --log-file gpurun_out/launches_infer.csv python scripts/profile_step.py 512 dense > gpurun_out/ncu_infer.log 2>&1
echo "ncu infer exit=$?"; wc -l gpurun_out/launches_infer.csv
python scripts/ncu_agg.py gpurun_out/launches_train.csv | head -40
python scripts/ncu_agg.py gpurun_out/launches_infer.csv | head -30
