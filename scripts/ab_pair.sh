#!/bin/bash
# A/B on one box: the library of the last commit (build_ab/) vs the working tree with SC_GEMM_MULTICAST = 0 / 2.
# needs build_ab/:  mkdir build_ab && git archive HEAD sparse-image-captioning_b200/csrc include | tar -x -C build_ab &&
#                   make -C build_ab/sparse-image-captioning_b200/csrc -j8        (delete build_ab/ afterwards)
L=sparse-image-captioning_b200/csrc/libsc_b200.so
cp $L /tmp/new.so
run() { python bench.py --steps 16 --warmup 8 --no-cpu-baseline 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$1', 'infer', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'train', round(d['train']['ms_per_step'],3))"; }
cp build_ab/$L $L; run head
cp /tmp/new.so $L; SC_GEMM_MULTICAST=0 run new_mc0
SC_GEMM_MULTICAST=2 run new_mc2
cp build_ab/$L $L; run head
cp /tmp/new.so $L; SC_GEMM_MULTICAST=2 run new_mc2
