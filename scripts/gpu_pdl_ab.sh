#!/bin/bash
for rep in 1 2; do for m in 0 1 2 3; do echo "mask=$m: $(SC_WALL_ONLY=1 SC_PDL_MASK=$m python scripts/profile_train.py 2>&1 | tail -1)"; done; done
