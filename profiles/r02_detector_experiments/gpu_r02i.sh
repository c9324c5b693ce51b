#!/bin/bash
# Round 2, pass i: pooled-tail k_detector against the previous one (GPET_DET_POOLED=0): GPU suite with each, kernel times,
# configs 4 / 5, ncu --set full of the source.txt frame with the pooled kernel.
TAG=${1:-r02i}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest (pooled)"; timeout -k 10 1200 python -m pytest tests -m gpu -q --tb=short -x > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -12 $OUT/${TAG}_pytest_gpu.log
for v in 0 1; do
  GPET_DET_POOLED=$v python tools/kprof.py --source source.txt --reps 20 --flush > $OUT/${TAG}_kprof_source_pooled$v.txt 2>&1
  GPET_DET_POOLED=$v python tools/kprof.py --source pointsource.txt --reps 20 --flush > $OUT/${TAG}_kprof_point_pooled$v.txt 2>&1
  GPET_DET_POOLED=$v python tools/kprof_configs.py > $OUT/${TAG}_kprof_configs_pooled$v.txt 2>&1
  GPET_DET_POOLED=$v python tools/bigframes_sweep.py --scales 1,4 > $OUT/${TAG}_bigframes_pooled$v.txt 2>&1
  echo "-- pooled=$v"; cat $OUT/${TAG}_kprof_source_pooled$v.txt | head -14; cat $OUT/${TAG}_bigframes_pooled$v.txt; grep -i "k_detector\|pairs/s" $OUT/${TAG}_kprof_configs_pooled$v.txt | head
done
echo "== ncu"
timeout -k 10 900 ncu --set full --clock-control none --import-source on --launch-skip 27 -c 9 -f -o $OUT/${TAG}_frame_full python tools/kprof.py --source source.txt --reps 2 > $OUT/${TAG}_ncu_frame.log 2>&1
ls -la $OUT/${TAG}*
