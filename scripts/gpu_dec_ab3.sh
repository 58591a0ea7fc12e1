#!/bin/bash
# Decode-step A/B: dynamic batching (G queued 512-image batches decoded as one device batch) x pipeline slots.
source scripts/gpu_dec_ab_lib.sh
run base_s8 A=1 -- --slots 8
run g2_s4 A=1 -- --coalesce 2 --slots 4
run g4_s2 A=1 -- --coalesce 4 --slots 2
run g4_s3 A=1 -- --coalesce 4 --slots 3
run g8_s1 A=1 -- --coalesce 8 --slots 1
run g8_s2 A=1 -- --coalesce 8 --slots 2
run g12_s2 A=1 -- --coalesce 12 --slots 2
