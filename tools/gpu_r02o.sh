#!/bin/bash
# Round 2, pass o: A/B of the shared-memory interaction tables in k_front (GPET_SMEM_TABLES), ncu launch list and `--set full`
# capture of the source.txt frame at HEAD, default bench line.
TAG=${1:-r02o}
OUT=gpurun_out
mkdir -p $OUT
for v in 0 1; do
  GPET_SMEM_TABLES=$v python tools/kprof.py --source source.txt --reps 20 --flush > $OUT/${TAG}_kprof_source_smemtab$v.txt 2>&1
  GPET_SMEM_TABLES=$v python tools/kprof_configs.py > $OUT/${TAG}_kprof_configs_smemtab$v.txt 2>&1
  echo "-- smem tables=$v"; grep -i "k_front\|per frame" $OUT/${TAG}_kprof_source_smemtab$v.txt; grep -i "k_front\|pairs/s" $OUT/${TAG}_kprof_configs_smemtab$v.txt
  GPET_SMEM_TABLES=$v timeout -k 10 300 ncu --metrics gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:k_front --launch-skip 3 -c 2 --csv --log-file $OUT/${TAG}_ncu_kfront_smemtab$v.csv python tools/kprof.py --source source.txt --reps 2 > /dev/null 2>&1
  GPET_SMEM_TABLES=$v timeout -k 10 300 ncu --metrics gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:k_front --launch-skip 6 -c 6 --csv --log-file $OUT/${TAG}_ncu_kfront_configs_smemtab$v.csv python tools/kprof_configs.py --reps 1 > /dev/null 2>&1
done
echo "== bench"; timeout -k 10 900 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench exit $?"; python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["with_48_byte_records"], d["e2e_files"]["value"], d["roofline"]["frac"])
PY
echo "== ncu launch list of the bench command"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --frames-per-step 8 --e2e-frames-per-step 8 > $OUT/${TAG}_ncu_bench.log 2>&1
echo "== ncu full"
timeout -k 10 900 ncu --set full --clock-control none --import-source on --launch-skip 27 -c 9 -f -o $OUT/${TAG}_frame_full python tools/kprof.py --source source.txt --reps 2 > $OUT/${TAG}_ncu_frame.log 2>&1
ls -la $OUT/${TAG}*
