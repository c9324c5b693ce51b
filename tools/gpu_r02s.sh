#!/bin/bash
# Round 2, pass s: GPU suite, kernel times (k_prep with pipelined loads), default bench line.
TAG=${1:-r02s}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest"; timeout -k 10 900 python -m pytest tests -m gpu -q --tb=short -x --timeout 300 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -8 $OUT/${TAG}_pytest_gpu.log
timeout -k 5 120 python tools/kprof.py --source source.txt --reps 20 --flush > $OUT/${TAG}_kprof_source.txt 2>&1
timeout -k 5 120 python tools/bigframes_sweep.py --scales 1,4 > $OUT/${TAG}_bigframes.txt 2>&1
cat $OUT/${TAG}_kprof_source.txt; cat $OUT/${TAG}_bigframes.txt
echo "== bench"; timeout -k 10 900 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench exit $?"; python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["with_48_byte_records"]["value"], d["e2e_files"]["value"], d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline"]["us_per_launch"])
PY
