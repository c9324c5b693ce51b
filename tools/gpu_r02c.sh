#!/bin/bash
# Round 2, third GPU pass: the whole GPU suite (exchange / halo / 64-bit id / file-writer tests included), per-kernel times,
# the bench line of both arms.  Usage: gpurun --timeout 1800 -- bash tools/gpu_r02c.sh r02c
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest"; timeout -k 10 1200 python -m pytest tests -m gpu -q --tb=short -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; grep -E "passed|failed|frame cuts|Error|error" $OUT/${TAG}_pytest_gpu.log | tail -12
echo "== kprof"
timeout -k 10 300 python tools/kprof.py --source source.txt --reps 20 --flush > $OUT/${TAG}_kprof_source.txt 2>&1; cat $OUT/${TAG}_kprof_source.txt
echo "== bench"
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench exit $?"; tail -5 $OUT/${TAG}_bench_n1.err
python - <<PY
import json
try:
    j = json.load(open("$OUT/${TAG}_bench_n1.json"))
    print("value %.4g e2e %.4g e2e_files %s ms/step %.2f launches %d" % (j["value"], j["e2e"]["value"], j["e2e_files"] and "%.4g" % j["e2e_files"]["value"], j["ms_per_step"], j["gpu_launches"]))
    print("roofline", j["roofline"]["kernel"], j["roofline"]["frac"], "cpu_baseline", j["cpu_baseline"] and j["cpu_baseline"]["value"], "clocks", j["clocks"])
    print("extra", j["extra"])
except Exception as e:
    print("bench line unreadable:", e)
PY
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference_n1.json 2> $OUT/${TAG}_bench_reference_n1.err; echo "ref exit $?"; cut -c1-400 $OUT/${TAG}_bench_reference_n1.json
ls -la $OUT | tail -8
