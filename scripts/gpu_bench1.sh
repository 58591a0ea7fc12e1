#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "csr" > gpurun_out/k_csr.log 2>&1; echo "csr exit=$? $(tail -n 1 gpurun_out/k_csr.log)"
timeout -s KILL 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -n 2 gpurun_out/smoke.log)"
timeout -s KILL 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err; echo "bench exit=$?"; tail -c 3000 gpurun_out/bench_dense.json; tail -n 5 gpurun_out/bench_dense.err
timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --backend csr --no-cpu-baseline > gpurun_out/bench_csr.json 2> gpurun_out/bench_csr.err; echo "bench csr exit=$?"; tail -c 2500 gpurun_out/bench_csr.json; tail -n 5 gpurun_out/bench_csr.err
