#!/bin/bash
# final launch lists of this round (profiler range: exactly one eager training step / one un-graphed inference step)
mkdir -p gpurun_out
SC_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python scripts/profile_train.py > gpurun_out/ncu_train.log 2>&1
SC_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infer.csv python scripts/profile_step.py 512 dense > gpurun_out/ncu_infer.log 2>&1
python scripts/ncu_agg.py gpurun_out/launches_train.csv 45 | tee gpurun_out/launches_train_summary.txt
python scripts/ncu_agg.py gpurun_out/launches_infer.csv 30 | tee gpurun_out/launches_infer_summary.txt
SC_NCU_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:sc_gemm_bf16_kernel<.int.256, .bool.0, .int.3, .int.3" -s 3 -c 2 -f -o gpurun_out/inf_gemm_topk python scripts/profile_step.py 512 dense > gpurun_out/ncu_inf_topk.log 2>&1
python scripts/ncu_metrics.py gpurun_out/inf_gemm_topk.ncu-rep | tee gpurun_out/inf_gemm_topk_summary.txt
rm -f gpurun_out/*.ncu-rep
