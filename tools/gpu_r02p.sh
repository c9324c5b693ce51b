#!/bin/bash
# Round 2, pass p: k_detector at 4 blocks x 256 threads per SM (64 registers, 16 staging rows per warp): GPU suite, kernel times,
# ncu --set full of the detector kernel.
TAG=${1:-r02p}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest"; timeout -k 10 900 python -m pytest tests -m gpu -q --tb=short -x --timeout 300 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -8 $OUT/${TAG}_pytest_gpu.log
timeout -k 5 120 python tools/kprof.py --source source.txt --reps 20 --flush > $OUT/${TAG}_kprof_source.txt 2>&1
timeout -k 5 120 python tools/kprof.py --source pointsource.txt --reps 20 --flush > $OUT/${TAG}_kprof_point.txt 2>&1
timeout -k 5 120 python tools/kprof_configs.py > $OUT/${TAG}_kprof_configs.txt 2>&1
timeout -k 5 120 python tools/bigframes_sweep.py --scales 1,4 > $OUT/${TAG}_bigframes.txt 2>&1
cat $OUT/${TAG}_kprof_source.txt; grep -i "k_detector" $OUT/${TAG}_kprof_point.txt; cat $OUT/${TAG}_bigframes.txt; grep -i "k_detector\|pairs/s" $OUT/${TAG}_kprof_configs.txt | cut -c1-120
echo "== ncu"
timeout -k 10 900 ncu --set full --clock-control none --import-source on --launch-skip 28 -c 1 -f -o $OUT/${TAG}_det_full python tools/kprof.py --source source.txt --reps 2 > $OUT/${TAG}_ncu_det.log 2>&1
ls -la $OUT/${TAG}*
