#!/bin/bash
# --set full captures (3 launches each) of the kernels that lead the inference launch list (un-graphed step, profiler range)
mkdir -p gpurun_out
for spec in "gemm64|regex:sc_gemm_bf16_kernel<64|200|3" "gemm128|regex:sc_gemm_bf16_kernel<128|40|3" "gemm256|regex:sc_gemm_bf16_kernel<256|3|3" "xattn|regex:cross_attn_mma_kernel|20|2" "selfattn|regex:self_attn_step_kernel|40|2" "beam|regex:beam_row_kernel|4|2" "ln|regex:layernorm_kernel|60|2"; do
  IFS='|' read -r name pat skip cnt <<< "$spec"
  SC_NCU_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k "$pat" -s $skip -c $cnt -f -o gpurun_out/inf_$name python scripts/profile_step.py 512 dense > gpurun_out/ncu_inf_$name.log 2>&1
  echo "$name exit=$?"
done
for spec in "tgemm_fwd|regex:sc_gemm_bf16_kernel<128, 0, 3, 1, 0>|4|2" "tgemm_wgrad|regex:sc_gemm_bf16_kernel<128, 0, 3, 0, 1>|4|2" "tattn_bwd|regex:attn_train_bwd_mma_kernel|2|2" "tattn_fwd|regex:attn_train_fwd_mma_kernel|2|2" "tadam|regex:adam_clip_kernel|0|2" "tmaskb|regex:apply_mask_batched_kernel|0|1"; do
  IFS='|' read -r name pat skip cnt <<< "$spec"
  SC_NCU_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k "$pat" -s $skip -c $cnt -f -o gpurun_out/tr_$name python scripts/profile_train.py > gpurun_out/ncu_tr_$name.log 2>&1
  echo "$name exit=$?"
done
for f in gpurun_out/inf_*.ncu-rep gpurun_out/tr_t*.ncu-rep; do echo "## $f"; python scripts/ncu_metrics.py $f; done | tee gpurun_out/r01b_ncu_full_summary.txt
