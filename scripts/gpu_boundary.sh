#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_boundary_gpu.py tests/test_dropin_gpu.py -q -m gpu -s 2>&1 | tee gpurun_out/boundary.txt | tail -80
