#!/bin/bash
# Quick GPU check: parity tests + per-kernel times + bench line.  Usage: gpurun -- bash tools/gpu_quick.sh tag [pytest-k-expr]
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q ${2:+-k "$2"} > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
tail -15 $OUT/${TAG}_pytest_gpu.log
python tools/kprof.py --source source.txt --reps 20 --flush > $OUT/${TAG}_kprof_source.txt 2>&1
python tools/kprof.py --source pointsource.txt --reps 20 --flush > $OUT/${TAG}_kprof_point.txt 2>&1
cat $OUT/${TAG}_kprof_source.txt

python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench exit $?"
python -c "
import json; j=json.load(open('$OUT/${TAG}_bench_n1.json')); print('value', j['value'], 'e2e', j['e2e']['value'], 'ms', j['ms_per_step'])"
