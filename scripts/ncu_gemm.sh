#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:sc_gemm_bf16 -s 60 -c 8 -f -o gpurun_out/prof_gemm_dec python scripts/profile_step.py 512 dense > gpurun_out/ncu_gemm_dec.log 2>&1; echo "dec exit=$?"
ncu --set full --clock-control none --import-source on -k regex:sc_gemm_bf16 -s 0 -c 8 -f -o gpurun_out/prof_gemm_enc python scripts/profile_step.py 512 dense > gpurun_out/ncu_gemm_enc.log 2>&1; echo "enc exit=$?"
tail -3 gpurun_out/ncu_gemm_dec.log
ls -la gpurun_out/*.ncu-rep
