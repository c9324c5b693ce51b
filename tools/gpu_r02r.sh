#!/bin/bash
# Round 2, 8-GPU pass r: the driver's own bench command at N = 8 (per-rank step times in details), then N = 4 and N = 1.
TAG=${1:-r02r}
OUT=gpurun_out
mkdir -p $OUT
for n in 8 4; do
echo "== bench N=$n"; timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > $OUT/${TAG}_bench_n$n.json 2> $OUT/${TAG}_bench_n$n.err; echo "bench exit $?"; tail -2 $OUT/${TAG}_bench_n$n.err
python - <<PY
import json
try:
    j = json.loads(open("$OUT/${TAG}_bench_n$n.json").read().strip().splitlines()[-1])
    print("N", j["n_gpus"], "value %.4g e2e %.4g (48B %.4g) ms/step %.2f" % (j["value"], j["e2e"]["value"], j["e2e"]["with_48_byte_records"]["value"], j["ms_per_step"]))
    print(j["details"])
except Exception as e:
    print("bench line unreadable:", e)
PY
done
echo "== bench N=1"; timeout -k 10 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; python -c "
import json; j=json.loads(open('$OUT/${TAG}_bench_n1.json').read().strip().splitlines()[-1]); print('N 1 value %.4g e2e %.4g files %.4g' % (j['value'], j['e2e']['value'], j['e2e_files']['value'])); print(j['details'])"
