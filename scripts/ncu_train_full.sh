#!/bin/bash
# --set full captures (2 launches each) of the non-GEMM kernels that lead the training launch list
mkdir -p gpurun_out
for spec in "attn_bwd|regex:attn_bwd_kernel|2" "attn_fwd|regex:attn_fwd_kernel|2" "boxb|regex:box_bias_bwd_kernel|1" "lnb|regex:layernorm_bwd_kernel|2" "prep|regex:prep_grad_kernel|2" "maskt|regex:apply_mask_t_kernel|2" "mgr|regex:mask_grad_reduce_kernel|2"; do
  IFS='|' read -r name pat cnt <<< "$spec"
  SC_NCU_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k "$pat" -c $cnt -f -o gpurun_out/tr_$name python scripts/profile_train.py > gpurun_out/ncu_tr_$name.log 2>&1
  echo "$name exit=$?"
done
for f in gpurun_out/tr_*.ncu-rep; do echo "## $f"; python scripts/ncu_metrics.py $f; done | tee gpurun_out/tr_full_summary.txt
