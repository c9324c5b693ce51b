#!/bin/bash
# Round 2, multi-GPU pass (run with `gpurun --gpus N`): NCCL paths of bench.py (frame-sharded acquisition, exchange arm) and
# the named-scale acquisitions.  Usage: gpurun --gpus 2 --timeout 1500 -- bash tools/gpu_r02d.sh r02d 2 1e8 1e9
TAG=${1:-r02d}; N=${2:-2}; D4=${3:-1e9}; D5=${4:-1e10}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpus.txt
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== gpu tests needing one GPU (exchange)"; timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "exchange or halo or history or file_run or frame_cuts" > $OUT/${TAG}_pytest_subset.log 2>&1; tail -3 $OUT/${TAG}_pytest_subset.log
echo "== bench N=$N"; timeout -k 10 900 $RUN bench.py --gpus $N --steps 10 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench exit $?"; tail -3 $OUT/${TAG}_bench_n$N.err
python - <<PY
import json
try:
    j = json.load(open("$OUT/${TAG}_bench_n$N.json"))
    print("N", j["n_gpus"], "value %.4g e2e %.4g ms/step %.2f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]), "exchange", j["exchange"])
    print("counters", j["counters"])
except Exception as e:
    print("bench line unreadable:", e)
PY
echo "== bench N=1 (same box)"; timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; python -c "
import json; j=json.load(open('$OUT/${TAG}_bench_n1.json')); print('N 1 value %.4g e2e %.4g' % (j['value'], j['e2e']['value']))"
echo "== config 4 at $D4 decays on $N GPUs"; timeout -k 10 900 $RUN tools/scale_runs.py --config config4_mouse --decays $D4 > $OUT/${TAG}_config4_n$N.json 2> $OUT/${TAG}_config4_n$N.err; echo "exit $?"; cut -c1-700 $OUT/${TAG}_config4_n$N.json; tail -2 $OUT/${TAG}_config4_n$N.err
echo "== config 5 at $D5 decays on $N GPUs"; timeout -k 10 1200 $RUN tools/scale_runs.py --config config5_ring --decays $D5 > $OUT/${TAG}_config5_n$N.json 2> $OUT/${TAG}_config5_n$N.err; echo "exit $?"; cut -c1-700 $OUT/${TAG}_config5_n$N.json; tail -2 $OUT/${TAG}_config5_n$N.err
ls -la $OUT | tail -8
