#!/bin/bash
# Round 2, pass v: digitizer grid sizes (GPET_DIGI_GRID = blocks per SM of every digitizer kernel) -- fewer resident tiles, less polling?
TAG=${1:-r02v}
OUT=gpurun_out
mkdir -p $OUT
for g in 2 3; do
GPET_DIGI_GRID=$g timeout -k 5 120 python tools/kprof.py --source source.txt --reps 20 --flush > $OUT/${TAG}_kprof_source_grid$g.txt 2>&1
GPET_DIGI_GRID=$g timeout -k 5 120 python tools/bigframes_sweep.py --scales 1,4 > $OUT/${TAG}_bigframes_grid$g.txt 2>&1
echo "-- grid $g"; cat $OUT/${TAG}_kprof_source_grid$g.txt | grep -v k_front; cat $OUT/${TAG}_bigframes_grid$g.txt
done
