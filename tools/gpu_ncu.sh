#!/bin/bash
# ncu --set full of selected kernels of one warm source.txt frame.  Usage: gpurun -- bash tools/gpu_ncu.sh tag 'regex' [count]
TAG=${1:-n}; RE=${2:-k_detector}; CNT=${3:-1}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$RE" --launch-skip 3 -c $CNT -f -o gpurun_out/${TAG} python tools/kprof.py --source source.txt --reps 2 > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log; ls -la gpurun_out/${TAG}.ncu-rep
