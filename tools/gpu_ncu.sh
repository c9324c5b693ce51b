#!/bin/bash
# ncu --set full capture of one warm source.txt frame (9 kernels).  Usage: gpurun -- bash tools/gpu_ncu.sh tag
TAG=${1:-ncu}
OUT=gpurun_out
mkdir -p $OUT
ncu --set full --clock-control none --import-source on --launch-skip 27 -c 9 -f -o $OUT/${TAG}_frame_full python tools/kprof.py --source source.txt --reps 2 > $OUT/${TAG}_ncu_frame.log 2>&1
ls -la $OUT
