#!/bin/bash
# SMP training step vs the split count of the weight-gradient GEMMs on the side stream (fewer splits = fewer CTAs beside the main chain)
for sp in 4 2 1 3; do
  echo "SC_WGRAD_MAX_SPLITS=$sp: $(SC_WGRAD_MAX_SPLITS=$sp SC_WALL_ONLY=1 python scripts/profile_train.py 2>&1 | tail -1)"
done
