#!/bin/bash
# Sweep the launch-shape knobs of the transport kernels (per-kernel CUDA-event times).  Usage: gpurun -- bash tools/gpu_sweep.sh tag
TAG=${1:-sw}
OUT=gpurun_out/${TAG}_sweep.txt
mkdir -p gpurun_out; : > $OUT
python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -12 gpurun_out/${TAG}_pytest_gpu.log
for r in 1 2 4 8 12 16; do
  echo "== GPET_REFILL_MIN=$r" >> $OUT
  GPET_REFILL_MIN=$r python tools/kprof.py --source source.txt --reps 20 --flush 2>&1 | grep -E "k_detector|k_front" >> $OUT
done
for g in 1 4 8 12 16 24 32; do for e in 1 8 16 24; do
  echo "== GPET_GEN_MIN=$g GPET_ENTRY_MIN=$e" >> $OUT
  GPET_GEN_MIN=$g GPET_ENTRY_MIN=$e python tools/kprof.py --source source.txt --reps 20 --flush 2>&1 | grep -E "k_front" >> $OUT
done; done
cat $OUT
