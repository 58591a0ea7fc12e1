#!/bin/bash
# Parity at the BASELINE configurations (tests/test_parity_baseline_gpu.py, printed numbers kept) + smoke().
mkdir -p gpurun_out
python -m pytest tests/test_parity_baseline_gpu.py -q -m gpu -s -x 2>&1 | tee gpurun_out/parity_baseline.txt | tail -40
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.txt | tail -3
