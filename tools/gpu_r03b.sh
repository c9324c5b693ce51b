#!/bin/bash
# Round 2, last 8-GPU pass: the bench at N = 8 and N = 1 on the same box at HEAD (frames of ~8 M pairs).
TAG=${1:-r03b}
OUT=gpurun_out
mkdir -p $OUT
echo "== bench N=8"; timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n8.json 2> $OUT/${TAG}_bench_n8.err; echo "bench exit $?"; tail -2 $OUT/${TAG}_bench_n8.err
echo "== bench N=1"; timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
python - <<PY
import json
for n in (8, 1):
    try:
        j = json.loads(open("$OUT/${TAG}_bench_n%d.json" % n).read().strip().splitlines()[-1])
        print("N", j["n_gpus"], "value %.4g e2e %.4g ms/step %.2f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]), j["details"]["frames_per_step_per_gpu"], j["details"]["mean_step_ms_by_rank"], j.get("exchange"))
    except Exception as e:
        print("bench line unreadable:", n, e)
PY
