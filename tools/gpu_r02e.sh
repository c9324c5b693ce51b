#!/bin/bash
# Round 2, single-GPU pass after the k_front trims and the regenerated tables: GPU suite, per-kernel times, config 4 fixture.
TAG=${1:-r02e}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest"; timeout -k 10 1200 python -m pytest tests -m gpu -q --tb=short > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -6 $OUT/${TAG}_pytest_gpu.log
echo "== kprof"
timeout -k 10 300 python tools/kprof.py --source source.txt --reps 20 --flush > $OUT/${TAG}_kprof_source.txt 2>&1; cat $OUT/${TAG}_kprof_source.txt
timeout -k 10 600 python tools/kprof_configs.py --decays 4000000 > $OUT/${TAG}_kprof_configs.txt 2>&1; grep "^#\|k_front\|k_detector" $OUT/${TAG}_kprof_configs.txt
echo "== reference statistics (config 4 with the mixed bone tables)"; timeout -k 10 600 python tools/ref_stats.py --configs config4_mouse > $OUT/${TAG}_ref_stats.log 2>&1; echo "ref_stats exit $?"; grep -v "^+" $OUT/${TAG}_ref_stats.log | tail -20
