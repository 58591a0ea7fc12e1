#!/bin/bash
# Decode-step A/B: tiles per persistent GEMM CTA in the 8-batches-in-flight regime.
source scripts/gpu_dec_ab_lib.sh
T=qkv=3256,o=3256,cq=3256,co=3256,ff1=3256,ff2=3256
run all256_8 SC_DEC_TILES=$T -- --slots 8
run tpc2 SC_GEMM_TPC=2 SC_DEC_TILES=$T -- --slots 8
run tpc3 SC_GEMM_TPC=3 SC_DEC_TILES=$T -- --slots 8
run tpc2_128 SC_GEMM_TPC=2 SC_DEC_TILES=qkv=3128,o=3128,cq=3128,co=3128,ff1=3128,ff2=3128 -- --slots 8
run tpc2_s6 SC_GEMM_TPC=2 SC_DEC_TILES=$T -- --slots 6
run tpc2_s12 SC_GEMM_TPC=2 SC_DEC_TILES=$T -- --slots 12
