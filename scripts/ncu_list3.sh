#!/bin/bash
# launch lists (cold-cache, serialised: compare SHARES) of exactly one eager training step and one un-graphed inference
# step (profiler range), then an inference bench matrix (batches in flight x LayerNorm folding)
mkdir -p gpurun_out
SC_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python scripts/profile_train.py > gpurun_out/ncu_train.log 2>&1
echo "ncu train exit=$?"; wc -l gpurun_out/launches_train.csv
SC_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infer.csv python scripts/profile_step.py 512 dense > gpurun_out/ncu_infer.log 2>&1
echo "ncu infer exit=$?"; wc -l gpurun_out/launches_infer.csv
python scripts/ncu_agg.py gpurun_out/launches_train.csv 60 | tee gpurun_out/launches_train_summary.txt
python scripts/ncu_agg.py gpurun_out/launches_infer.csv 40 | tee gpurun_out/launches_infer_summary.txt
timeout -s KILL 300 python scripts/profile_train.py > gpurun_out/profile_train.txt 2>&1; head -50 gpurun_out/profile_train.txt
for flags in "" "--ln-fold"; do
for s in 2 4 6 8; do
  timeout -s KILL 300 python bench.py --steps 16 --warmup 8 --slots $s --no-train --no-cpu-baseline $flags > gpurun_out/bm.json 2> gpurun_out/bm.err
  python -c "import json;d=json.load(open('gpurun_out/bm.json'));print('$flags', $s, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))"
done; done
