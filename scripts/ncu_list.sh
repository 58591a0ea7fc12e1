#!/bin/bash
# launch list of one un-graphed inference step (cold-cache, serialised: compare SHARES)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1300 --csv --log-file gpurun_out/launches_r1.csv python scripts/profile_step.py ${1:-512} ${2:-dense} > gpurun_out/ncu_list.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/ncu_list.log; wc -l gpurun_out/launches_r1.csv
