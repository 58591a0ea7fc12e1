#!/bin/bash
# one --set full capture per kernel family (3 launches each) from an un-graphed inference step
mkdir -p gpurun_out
for spec in "gemm64|regex:sc_gemm_bf16_kernel<64|700" "gemm128|regex:sc_gemm_bf16_kernel<128|20" "box|regex:box_attention_kernel|2" "xattn|regex:cross_attn_step_kernel|20" "beam|regex:beam_step_kernel|4"; do
  IFS='|' read -r name pat skip <<< "$spec"
  ncu --set full --clock-control none --import-source on -k "$pat" -s $skip -c 3 -f -o gpurun_out/prof_$name python scripts/profile_step.py ${1:-512} dense > gpurun_out/ncu_$name.log 2>&1
  echo "$name exit=$?"
done
ls -la gpurun_out/*.ncu-rep
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -q -m gpu -k "decode_attention or beam_step or engine" 2>&1 | tail -3
