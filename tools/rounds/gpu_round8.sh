#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"; tail -3 gpurun_out/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tools/train_shards.py --N_rand 20 --steps 60 --hard_mul 2 > gpurun_out/train_shards_n$N.log 2>&1; echo "train rc=$?"; tail -2 gpurun_out/train_shards_n$N.log
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print(d.get("n_gpus"), "value", d.get("value"), "ms", d.get("ms_per_step"), "e2e", d.get("e2e", {}).get("value"), d.get("clocks"))
PY
