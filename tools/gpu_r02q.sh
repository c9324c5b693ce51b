#!/bin/bash
# Round 2, pass q: k_detector with the trimmed staging path (4 blocks x 256 per SM); A/B of k_front at 5 blocks per SM (48 registers).
TAG=${1:-r02q}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest"; timeout -k 10 900 python -m pytest tests -m gpu -q --tb=short -x --timeout 300 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -8 $OUT/${TAG}_pytest_gpu.log
for v in 4 5; do
GPET_FRONT_BLOCKS=$v timeout -k 5 120 python tools/kprof.py --source source.txt --reps 20 --flush > $OUT/${TAG}_kprof_source_front$v.txt 2>&1
GPET_FRONT_BLOCKS=$v timeout -k 5 120 python tools/kprof_configs.py > $OUT/${TAG}_kprof_configs_front$v.txt 2>&1
GPET_FRONT_BLOCKS=$v timeout -k 5 120 python tools/bigframes_sweep.py --scales 1,4 > $OUT/${TAG}_bigframes_front$v.txt 2>&1
echo "-- front blocks $v"; grep "k_front\|k_detector\|per frame" $OUT/${TAG}_kprof_source_front$v.txt; cat $OUT/${TAG}_bigframes_front$v.txt; grep -i "k_front\|k_detector\|pairs/s" $OUT/${TAG}_kprof_configs_front$v.txt | cut -c1-120
done
echo "== ncu"
timeout -k 10 900 ncu --set full --clock-control none --import-source on --launch-skip 27 -c 9 -f -o $OUT/${TAG}_frame_full python tools/kprof.py --source source.txt --reps 2 > $OUT/${TAG}_ncu_frame.log 2>&1
ls -la $OUT/${TAG}*
