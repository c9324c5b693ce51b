#!/bin/bash
# First GPU pass of a new round: what round 1 could not run any more (last bullet), then the usual pass.
# Usage: gpurun --timeout 600 -- bash tools/gpu_next_round.sh r02a
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
timeout 120 python tools/gpu_fuzz_digitizer.py --trials 3000 > $OUT/${TAG}_fuzz_digitizer.txt 2>&1; echo "fuzz exit $?"; tail -3 $OUT/${TAG}_fuzz_digitizer.txt
python -m pytest tests -m gpu -q --tb=short > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -3 $OUT/${TAG}_pytest_gpu.log
python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference_n1.json 2> $OUT/${TAG}_bench_reference_n1.err; echo "ref exit $?"
python tools/kprof.py --source source.txt --reps 20 --flush > $OUT/${TAG}_kprof_source.txt 2>&1; cat $OUT/${TAG}_kprof_source.txt
python tools/kprof_configs.py --decays 4000000 > $OUT/${TAG}_kprof_configs.txt 2>&1; grep "^#" $OUT/${TAG}_kprof_configs.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_source.csv python tools/kprof.py --source source.txt --reps 3 > $OUT/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 27 -c 9 -f -o $OUT/${TAG}_frame_full python tools/kprof.py --source source.txt --reps 2 > $OUT/${TAG}_ncu_frame.log 2>&1
ls -la $OUT | tail -12
