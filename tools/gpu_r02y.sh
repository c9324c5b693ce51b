#!/bin/bash
# Round 2, final pass: smoke(), GPU suite, default bench line, reference arm, ncu launch list of the bench command, ncu --set full
# of the source.txt frame at HEAD.
TAG=${1:-r02y}
OUT=gpurun_out
mkdir -p $OUT
echo "== smoke"; timeout -k 10 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== pytest"; timeout -k 10 1200 python -m pytest tests -m gpu -q --tb=short > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -4 $OUT/${TAG}_pytest_gpu.log
echo "== bench"; timeout -k 10 900 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench exit $?"; python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["with_48_byte_records"]["value"], d["e2e_files"]["value"], d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline"]["us_per_launch"], (d["cpu_baseline"] or {}).get("value"))
PY
echo "== reference arm"; timeout -k 10 900 python bench.py --impl reference > $OUT/${TAG}_bench_reference_n1.json 2> $OUT/${TAG}_bench_reference_n1.err; echo "ref exit $?"; python -c "
import json; j=json.loads(open('$OUT/${TAG}_bench_reference_n1.json').read().strip().splitlines()[-1]); print('reference', j['value'], j['ms_per_step'], j['steps'], j['warmup'])"
echo "== ncu launch list of the bench command"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --frames-per-step 8 --e2e-frames-per-step 8 > $OUT/${TAG}_ncu_bench.log 2>&1
echo "== ncu full"
timeout -k 10 900 ncu --set full --clock-control none --import-source on --launch-skip 27 -c 9 -f -o $OUT/${TAG}_frame_full python tools/kprof.py --source source.txt --reps 2 > $OUT/${TAG}_ncu_frame.log 2>&1
ls -la $OUT/${TAG}*
