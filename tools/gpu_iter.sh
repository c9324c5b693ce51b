#!/bin/bash
# Fast perf iteration: transport/digitizer parity subset + per-kernel times.  Usage: gpurun -- bash tools/gpu_iter.sh tag [pytest -k expr]
TAG=${1:-it}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q ${2:+-k "$2"} > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
tail -4 $OUT/${TAG}_pytest_gpu.log
python tools/kprof.py --source source.txt --reps 20 --flush > $OUT/${TAG}_kprof_source.txt 2>&1
python tools/kprof.py --source pointsource.txt --reps 20 --flush > $OUT/${TAG}_kprof_point.txt 2>&1
cat $OUT/${TAG}_kprof_source.txt; head -4 $OUT/${TAG}_kprof_point.txt
