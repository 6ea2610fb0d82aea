#!/bin/bash
mkdir -p gpurun_out/r2_16; O=gpurun_out/r2_16
N=${1:-8}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/gpu_ddp_check.py > $O/ddp_check_n$N.log 2>&1; echo "ddp rc=$?"; grep -E "world|Error|error" $O/ddp_check_n$N.log | tail -12
timeout 120 python bench.py --steps 50 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench1 rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 50 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench rc=$?"; tail -2 $O/bench_n$N.err
python - <<PY
import json
for f in ("bench_n1", "bench_n$N"):
    try:
        d = json.loads(open("$O/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d.get("n_gpus"), "value", d.get("value"), "ms", d.get("ms_per_step"), "e2e", d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
