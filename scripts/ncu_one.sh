#!/bin/bash
mkdir -p gpurun_out
for mode in plain ln; do
ncu --set full --clock-control none --import-source on -k regex:sc_gemm_bf16 -s 3 -c 1 -f -o gpurun_out/one_$mode python scripts/one_gemm.py 1536 2048 512 $mode > gpurun_out/ncu_one_$mode.log 2>&1; echo "$mode exit=$?"
done
ls -la gpurun_out/one_*.ncu-rep
