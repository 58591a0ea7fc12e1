#!/bin/bash
mkdir -p gpurun_out
SC_NCU_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:sc_gemm_bf16_kernel<.int.256, .bool.0, .int.3, .int.5" -s 3 -c 2 -f -o gpurun_out/inf_gemm_topk3 python scripts/profile_step.py 512 dense > gpurun_out/ncu_a.log 2>&1
SC_NCU_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:self_attn_step_kernel" -s 90 -c 2 -f -o gpurun_out/inf_selfattn2 python scripts/profile_step.py 512 dense > gpurun_out/ncu_b.log 2>&1
for f in gpurun_out/inf_gemm_topk3.ncu-rep gpurun_out/inf_selfattn2.ncu-rep; do echo "## $f"; python scripts/ncu_metrics.py $f; done | tee gpurun_out/r01c_ncu_extra.txt
rm -f gpurun_out/*.ncu-rep
