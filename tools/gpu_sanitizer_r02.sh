#!/bin/bash
# compute-sanitizer over a subset of the GPU parity tests at HEAD of round 2 (k_detector at 4 blocks per SM with 16 staging rows,
# k_pack_singles, compact delivery, exchange / emit window).
OUT=gpurun_out
mkdir -p $OUT
SUB="bit_exact_sizes or shipped_example or detector_transport or compact_singles or fused_front or edge_cases or time_slice_exchange"
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout -k 10 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SUB" > $OUT/r02_sanitizer_$tool.log 2>&1
  echo "exit $?"; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" $OUT/r02_sanitizer_$tool.log | tail -3
done
