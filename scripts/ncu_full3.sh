#!/bin/bash
# --set full captures of the tcgen05 GEMM (kernel names with template arguments need --kernel-name-base demangled)
mkdir -p gpurun_out
for spec in "gemm64|sc_gemm_bf16_kernel<.int.64,|200|3" "gemm128|sc_gemm_bf16_kernel<.int.128,|40|3" "gemm256|sc_gemm_bf16_kernel<.int.256,|3|3"; do
  IFS='|' read -r name pat skip cnt <<< "$spec"
  SC_NCU_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$pat" -s $skip -c $cnt -f -o gpurun_out/inf_$name python scripts/profile_step.py 512 dense > gpurun_out/ncu_inf_$name.log 2>&1
  echo "$name exit=$?"
done
for spec in "tgemm_fwd|sc_gemm_bf16_kernel<.int.128, .bool.0, .int.3, .int.1, .bool.0>|4|2" "tgemm_wgrad|sc_gemm_bf16_kernel<.int.128, .bool.0, .int.3, .int.0, .bool.1>|4|2"; do
  IFS='|' read -r name pat skip cnt <<< "$spec"
  SC_NCU_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$pat" -s $skip -c $cnt -f -o gpurun_out/tr_$name python scripts/profile_train.py > gpurun_out/ncu_tr_$name.log 2>&1
  echo "$name exit=$?"
done
for f in gpurun_out/inf_gemm*.ncu-rep gpurun_out/tr_tgemm*.ncu-rep; do echo "## $f"; python scripts/ncu_metrics.py $f; done | tee gpurun_out/r01b_ncu_gemm_summary.txt
# then the training checks of this round's row-kernel changes
bash scripts/gpu_train_check3.sh
# SASS evidence: tcgen05 / TMA mnemonics in the shipped library
cuobjdump -sass sparse-image-captioning_b200/csrc/libsc_b200.so 2>/dev/null | grep -oE "UTCHMMA|UTMALDG|UTCBAR|TCGEN05[A-Z.]*|UTCMMA|LDTM|STTM|UTMAPF|SYNCS[.A-Z]*" | sort | uniq -c | tee gpurun_out/sass_mnemonics.txt
ls -la gpurun_out/*.ncu-rep
