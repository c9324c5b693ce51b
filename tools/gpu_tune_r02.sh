#!/bin/bash
# Round 2: refill thresholds of the persistent transport kernels at the current launch shapes (environment knobs, no rebuild)
OUT=gpurun_out
mkdir -p $OUT
for v in 1 2 4 6 8 12; do
  echo "-- GPET_REFILL_MIN=$v"; GPET_REFILL_MIN=$v timeout -k 5 120 python tools/bigframes_sweep.py --scales 4 --reps 6 2>&1 | tail -1
done
for v in 4 8 12 16 20; do
  echo "-- GPET_GEN_MIN=$v GPET_ENTRY_MIN=$v"; GPET_GEN_MIN=$v GPET_ENTRY_MIN=$v timeout -k 5 120 python tools/bigframes_sweep.py --scales 4 --reps 6 2>&1 | tail -1
done
