#!/bin/bash
# Round 2, second GPU pass: parity + A/B of k_detector (cooperative Klein-Nishina sampling, staged rows, shared reciprocals)
# against round 1's kernel, the reference-binary transport statistics in their final layout, ncu of the frame.
# Usage: gpurun --timeout 1800 -- bash tools/gpu_r02b.sh r02b
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest"; timeout -k 10 1200 python -m pytest tests -m gpu -q --tb=short > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -8 $OUT/${TAG}_pytest_gpu.log
echo "== detector variants"
for v in 1 2; do
  GPET_DET_V=$v timeout -k 10 300 python tools/kprof.py --source source.txt --reps 20 --flush > $OUT/${TAG}_kprof_source_v$v.txt 2>&1
  grep -E "k_detector|k_front|per frame|counts" $OUT/${TAG}_kprof_source_v$v.txt
  GPET_DET_V=$v timeout -k 10 300 python tools/kprof.py --source pointsource.txt --reps 20 --flush > $OUT/${TAG}_kprof_point_v$v.txt 2>&1
  grep -E "k_detector|per frame" $OUT/${TAG}_kprof_point_v$v.txt
  GPET_DET_V=$v timeout -k 10 600 python tools/kprof_configs.py --decays 4000000 > $OUT/${TAG}_kprof_configs_v$v.txt 2>&1; grep "^#\|k_detector" $OUT/${TAG}_kprof_configs_v$v.txt
done
for rm in 1 8 12; do
  GPET_REFILL_MIN=$rm timeout -k 10 300 python tools/kprof.py --source source.txt --reps 10 --flush 2>&1 | grep -E "k_detector" | sed "s/^/refill_min=$rm /"
done
echo "== reference statistics"; timeout -k 10 1500 python tools/ref_stats.py > $OUT/${TAG}_ref_stats.log 2>&1; echo "ref_stats exit $?"; grep -v "^+" $OUT/${TAG}_ref_stats.log | tail -70
echo "== ncu"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_source.csv python tools/kprof.py --source source.txt --reps 3 > $OUT/${TAG}_ncu_launch.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on --launch-skip 27 -c 9 -f -o $OUT/${TAG}_frame_full python tools/kprof.py --source source.txt --reps 2 > $OUT/${TAG}_ncu_frame.log 2>&1
ls -la $OUT | tail -12
