#!/bin/bash
# ncu --set full of the gather SpMM (99.1 % and 95 % sparse, 1536 x 512 x 512) -> gpurun_out/gspmm.ncu-rep
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gspmm_kernel -o gpurun_out/gspmm -f python scripts/gspmm_once.py > gpurun_out/ncu_gspmm.log 2>&1
tail -3 gpurun_out/ncu_gspmm.log
