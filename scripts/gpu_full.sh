#!/bin/bash
# Full GPU check: all gpu tests, smoke, profile_step breakdown, default bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu -x > gpurun_out/tests.log 2>&1; echo "tests exit=$? $(tail -n 1 gpurun_out/tests.log)"
timeout -s KILL 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -n 1 gpurun_out/smoke.log)"
timeout -s KILL 300 python scripts/profile_step.py 512 dense > gpurun_out/profile_step.txt 2>&1; echo "profile exit=$?"; head -30 gpurun_out/profile_step.txt
timeout -s KILL 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -c 4000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
