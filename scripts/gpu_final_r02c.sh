#!/bin/bash
# end-of-session evidence on one box: GPU tier + smoke, every bench configuration + reference arms, ncu launch list of the bench command
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout -s KILL 1500 python -m pytest tests -q -m gpu -x > gpurun_out/tests.log 2>&1; echo "tests exit=$? $(tail -n 1 gpurun_out/tests.log)"
timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -1
bash scripts/gpu_bench_all.sh
bash scripts/ncu_list_r02.sh > /dev/null 2>&1; head -24 gpurun_out/r02_infer_launches_summary.txt
