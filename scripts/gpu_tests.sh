#!/bin/bash
# Whole GPU tier + smoke; printed parity numbers kept in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -s 2>&1 | tee gpurun_out/gpu_tests.txt | grep -E "^\[|passed|failed|Error|error|FAILED|assert" | tail -60
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.txt | tail -3
