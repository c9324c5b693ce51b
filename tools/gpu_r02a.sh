#!/bin/bash
# Round 2, first GPU pass: parity of the rewritten k_detector + A/B of its variants, reference-binary fixtures (digitizer
# pins with real dead-time kills, transport statistics of configs 1-thick / 2 / 4 / 5), digitizer fuzz, ncu evidence
# (full set of the frame, L2 access-policy window A/B on the big phantoms).
# Usage: gpurun --timeout 2400 -- bash tools/gpu_r02a.sh r02a
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpu.txt
echo "== pytest"; timeout -k 10 1200 python -m pytest tests -m gpu -q --tb=short -x > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -5 $OUT/${TAG}_pytest_gpu.log
echo "== detector variants"
for v in 1 2 3; do
  GPET_DET_V=$v timeout -k 10 300 python tools/kprof.py --source source.txt --reps 20 --flush > $OUT/${TAG}_kprof_source_v$v.txt 2>&1
  grep -E "k_detector|k_front|per frame|counts" $OUT/${TAG}_kprof_source_v$v.txt
done
for v in 1 2; do
  GPET_DET_V=$v timeout -k 10 300 python tools/kprof.py --source source.txt --reps 10 --flush --staged > $OUT/${TAG}_kprof_staged_v$v.txt 2>&1
done
GPET_DET_V=1 timeout -k 10 600 python tools/kprof_configs.py --decays 4000000 > $OUT/${TAG}_kprof_configs_v1.txt 2>&1; grep "^#\|k_detector\|k_front" $OUT/${TAG}_kprof_configs_v1.txt
timeout -k 10 600 python tools/kprof_configs.py --decays 4000000 > $OUT/${TAG}_kprof_configs_v2.txt 2>&1; grep "^#\|k_detector\|k_front" $OUT/${TAG}_kprof_configs_v2.txt
GPET_NO_L2_WINDOW=1 timeout -k 10 600 python tools/kprof_configs.py --decays 4000000 > $OUT/${TAG}_kprof_configs_nol2win.txt 2>&1; grep "^#\|k_front" $OUT/${TAG}_kprof_configs_nol2win.txt
echo "== fuzz"; timeout -k 10 200 python tools/gpu_fuzz_digitizer.py --trials 3000 > $OUT/${TAG}_fuzz_digitizer.txt 2>&1; echo "fuzz exit $?"; tail -3 $OUT/${TAG}_fuzz_digitizer.txt
echo "== reference pins"; timeout -k 10 900 python tools/ref_pin.py --no-transport > $OUT/${TAG}_ref_pin.log 2>&1; echo "ref_pin exit $?"; python - <<PY
import json
try:
    r = json.load(open("$OUT/ref_pin/report.json"))
    for c in r["cases"]:
        print(c["name"], c.get("adder_events"), c.get("ref_counts"), c.get("dead_time_kills_oracle"), c.get("oracle_vs_reference"), c.get("error", "")[:200])
except Exception as e:
    print("no report", e)
PY
echo "== reference statistics"; timeout -k 10 1500 python tools/ref_stats.py > $OUT/${TAG}_ref_stats.log 2>&1; echo "ref_stats exit $?"; grep -v "^+" $OUT/${TAG}_ref_stats.log | tail -80
echo "== ncu"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_source.csv python tools/kprof.py --source source.txt --reps 3 > $OUT/${TAG}_ncu_launch.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on --launch-skip 27 -c 9 -f -o $OUT/${TAG}_frame_full python tools/kprof.py --source source.txt --reps 2 > $OUT/${TAG}_ncu_frame.log 2>&1
# L2 access-policy window A/B on the 67 MB grids: k_front of config 5, with and without the window
for w in on off; do
  if [ $w = off ]; then export GPET_NO_L2_WINDOW=1; else unset GPET_NO_L2_WINDOW; fi
  timeout -k 10 600 ncu --metrics gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:k_front -s 4 -c 4 --csv --log-file $OUT/${TAG}_l2window_${w}_config5.csv python tools/kprof_configs.py --decays 4000000 --case config5 --reps 2 > $OUT/${TAG}_l2window_${w}.log 2>&1
  timeout -k 10 600 ncu --metrics gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:k_front -s 8 -c 8 --csv --log-file $OUT/${TAG}_l2window_${w}_config4.csv python tools/kprof_configs.py --decays 4000000 --case config4 --reps 2 > $OUT/${TAG}_l2window_${w}4.log 2>&1
done
unset GPET_NO_L2_WINDOW
ls -la $OUT | tail -30
