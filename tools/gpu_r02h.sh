#!/bin/bash
# Round 2, pass h: GPU suite at HEAD, default bench line (frames of ~4 M pairs), reference arm.
TAG=${1:-r02h}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest"; timeout -k 10 1200 python -m pytest tests -m gpu -q --tb=short > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -6 $OUT/${TAG}_pytest_gpu.log
echo "== bench"; timeout -k 10 900 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench exit $?"; python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["e2e_files"], d["roofline"]["frac"], d["roofline"]["us_per_launch"], (d["cpu_baseline"] or {}).get("value"))
PY
echo "== reference arm"; timeout -k 10 900 python bench.py --impl reference > $OUT/${TAG}_bench_reference_n1.json 2> $OUT/${TAG}_bench_reference_n1.err; echo "ref exit $?"; tail -c 600 $OUT/${TAG}_bench_reference_n1.json
