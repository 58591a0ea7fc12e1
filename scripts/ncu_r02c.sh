#!/bin/bash
# ncu --set full (source-level) of chosen kernels: SC_TARGETS=topk,o,cross,enc,box,ff2,self KREGEX=<kernel regex> -> gpurun_out/r02c_<tag>.ncu-rep
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
TAG=${TAG:-targets}
KREGEX=${KREGEX:-"sc_gemm_bf16_kernel|cross_attn_mma|enc_attn_mma|box_bias_all|self_attn_step"}
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k "regex:$KREGEX" \
  -o gpurun_out/r02c_$TAG -f python scripts/ncu_targets_r02c.py > gpurun_out/ncu_r02c.log 2>&1
tail -2 gpurun_out/ncu_r02c.log; ls -la gpurun_out/*.ncu-rep
python scripts/ncu_metrics.py gpurun_out/r02c_$TAG.ncu-rep > gpurun_out/r02c_${TAG}_summary.txt 2>&1; cat gpurun_out/r02c_${TAG}_summary.txt
