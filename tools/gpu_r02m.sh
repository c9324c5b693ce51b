#!/bin/bash
# Round 2, 8-GPU pass (gpurun --gpus 8): bench.py at N=8 (frame-sharded acquisition), config 4 at 1e9 decays on 8 GPUs,
# config 5 at 1e10 decays on 8 / 4 / 2 GPUs (BASELINE.json configs 4 / 5 at their named scale).
TAG=${1:-r02m}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpus.txt
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 "${@:2}"; }
echo "== bench N=8"; timeout -k 10 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline" > $OUT/${TAG}_bench_n8.json 2> $OUT/${TAG}_bench_n8.err; echo "bench exit $?"; tail -3 $OUT/${TAG}_bench_n8.err
python - <<PY
import json
try:
    j = json.loads(open("$OUT/${TAG}_bench_n8.json").read().strip().splitlines()[-1])
    print("N", j["n_gpus"], "value %.4g e2e %.4g ms/step %.2f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]))
except Exception as e:
    print("bench line unreadable:", e)
PY
echo "== bench N=1 (same box)"; timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; python -c "
import json; j=json.loads(open('$OUT/${TAG}_bench_n1.json').read().strip().splitlines()[-1]); print('N 1 value %.4g e2e %.4g' % (j['value'], j['e2e']['value']))"
echo "== config 4 at 1e9 decays on 8 GPUs"; timeout -k 10 600 bash -c "$(declare -f run); run 8 tools/scale_runs.py --config config4_mouse --decays 1e9" > $OUT/${TAG}_config4_n8.json 2> $OUT/${TAG}_config4_n8.err; echo "exit $?"; cut -c1-600 $OUT/${TAG}_config4_n8.json; tail -2 $OUT/${TAG}_config4_n8.err
for n in 8 4 2; do
echo "== config 5 at 1e10 decays on $n GPUs"; timeout -k 10 900 bash -c "$(declare -f run); run $n tools/scale_runs.py --config config5_ring --decays 1e10" > $OUT/${TAG}_config5_n$n.json 2> $OUT/${TAG}_config5_n$n.err; echo "exit $?"; cut -c1-600 $OUT/${TAG}_config5_n$n.json; tail -2 $OUT/${TAG}_config5_n$n.err
done
ls -la $OUT/${TAG}*
