#!/bin/bash
# ncu launch list (gpu__time_duration.sum, --clock-control none) of the default bench command's inference arm, first 6000 launches (two pipeline slots: the tile hints of the throughput regime are in force)
# (engine set-up + the first coalesced batches) -> profiles/r02_infer_launches{.csv,_summary.txt}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_infer_launches.csv \
    python bench.py --steps 10 --warmup 5 --no-train --no-cpu-baseline --coalesce 5 --slots 2 --e2e-coalesce 5 --e2e-slots 1 > gpurun_out/ncu_bench.log 2>&1
python scripts/ncu_agg.py gpurun_out/r02_infer_launches.csv 30 | tee gpurun_out/r02_infer_launches_summary.txt
