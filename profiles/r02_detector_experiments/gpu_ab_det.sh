#!/bin/bash
# A/B of the pooled-tail k_detector against the previous one (GPET_DET_SPLIT=0): GPU suite, kernel times, optional ncu.
TAG=${1:-ab}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest (split)"; timeout -k 10 600 python -m pytest tests -m gpu -q --tb=short -x --timeout 120 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -12 $OUT/${TAG}_pytest_gpu.log
for v in 0 1; do
  GPET_DET_SPLIT=$v timeout -k 5 120 python tools/kprof.py --source source.txt --reps 20 --flush > $OUT/${TAG}_kprof_source_split$v.txt 2>&1
  GPET_DET_SPLIT=$v timeout -k 5 120 python tools/kprof_configs.py > $OUT/${TAG}_kprof_configs_split$v.txt 2>&1
  GPET_DET_SPLIT=$v timeout -k 5 120 python tools/bigframes_sweep.py --scales 1,4 > $OUT/${TAG}_bigframes_split$v.txt 2>&1
  echo "-- split=$v"; grep -i "k_detector\|k_adder\|per frame" $OUT/${TAG}_kprof_source_split$v.txt; cat $OUT/${TAG}_bigframes_split$v.txt; grep -i "k_detector\|k_adder\|pairs/s" $OUT/${TAG}_kprof_configs_split$v.txt | head
done
if [ -n "$2" ]; then
echo "== ncu"
timeout -k 10 900 ncu --set full --clock-control none --import-source on --launch-skip 30 -c 3 -f -o $OUT/${TAG}_det_full python tools/kprof.py --source source.txt --reps 2 > $OUT/${TAG}_ncu_det.log 2>&1
fi
ls -la $OUT/${TAG}*
