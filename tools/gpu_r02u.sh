#!/bin/bash
# Round 2, pass u: A/B of the digitizer kernels at higher occupancy (k_emit_singles 4 blocks / SM at 64 registers, k_prep 5,
# k_bucket_scatter 6, k_coinc 5): the alternative library is swapped in for the second half.
TAG=${1:-r02u}
OUT=gpurun_out
mkdir -p $OUT
run() {
timeout -k 5 120 python tools/kprof.py --source source.txt --reps 20 --flush > $OUT/${TAG}_kprof_source_$1.txt 2>&1
timeout -k 5 120 python tools/bigframes_sweep.py --scales 1,4 > $OUT/${TAG}_bigframes_$1.txt 2>&1
echo "-- $1"; cat $OUT/${TAG}_kprof_source_$1.txt; cat $OUT/${TAG}_bigframes_$1.txt
}
run base
cp gpet_b200/libgpet_b200.so /tmp/base.so; cp gpet_b200/libgpet_b200_hi.so gpet_b200/libgpet_b200.so
run hi
timeout -k 10 600 python -m pytest tests -m gpu -q --tb=short -x --timeout 300 -k "digitiz or coinc or run_" > $OUT/${TAG}_pytest_hi.log 2>&1; tail -3 $OUT/${TAG}_pytest_hi.log
cp /tmp/base.so gpet_b200/libgpet_b200.so
