#!/bin/bash
# training-step diagnostics on one box: the whole step, the main chain alone (weight-gradient GEMMs skipped), one stream
echo "step: $(SC_WALL_ONLY=1 python scripts/profile_train.py 2>&1 | tail -1)"
echo "main chain alone (weight-gradient GEMMs skipped): $(SC_SKIP_WGRAD=1 SC_WALL_ONLY=1 python scripts/profile_train.py 2>&1 | tail -1)"
echo "one stream (SC_WGRAD_RING=1): $(SC_WGRAD_RING=1 SC_WALL_ONLY=1 python scripts/profile_train.py 2>&1 | tail -1)"
