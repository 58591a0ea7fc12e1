#!/bin/bash
# First-contact GPU run: each group in its own process under a hard timeout so that a hung kernel cannot take the
# whole call down.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1 to=$2; shift 2
  echo "=== $name ===" | tee -a gpurun_out/summary.txt
  timeout -s KILL $to "$@" > gpurun_out/$name.log 2>&1
  echo "exit=$? $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt
}
run k_fp32    300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "fp32 or layernorm or embed or binarize or reorder"
run k_tc_dense 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "bf16_dense"
run k_tc_mask 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "masked_prologue or bernoulli"
run k_attn    300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "box_attention or decode_attention"
run k_beam    300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "beam_step or csr"
run engine    600 python -m pytest tests/test_engine_gpu.py -q -m gpu
cat gpurun_out/summary.txt
