#!/bin/bash
mkdir -p gpurun_out
SC_TRAIN_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 40 -c 2 -f -o gpurun_out/attn_fwd python scripts/profile_train.py > gpurun_out/ncu_attn_fwd.log 2>&1; echo "fwd exit=$?"
SC_TRAIN_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 40 -c 2 -f -o gpurun_out/attn_bwd python scripts/profile_train.py > gpurun_out/ncu_attn_bwd.log 2>&1; echo "bwd exit=$?"
SC_TRAIN_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:layernorm_bwd_kernel -s 40 -c 1 -f -o gpurun_out/ln_bwd python scripts/profile_train.py > gpurun_out/ncu_ln_bwd.log 2>&1; echo "lnbwd exit=$?"
