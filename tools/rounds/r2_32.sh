#!/bin/bash
# round-2 two-GPU visit: config-5 sweep across 2 ranks, data-parallel training with the hard-ray pool, then (one GPU) the GPU suite
# after the pool-kernel change and the launch list of a pool-mode training run
O=gpurun_out/r2_32; mkdir -p $O
N=${1:-2}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/render_sweep.py > $O/render_sweep_n$N.log 2>&1; echo "render sweep rc=$?"; grep -E "^\{" $O/render_sweep_n$N.log | cut -c1-250
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/train_shards.py --N_rand 1 --hard_ratio 0.2 --hard_mul 2 --steps 300 > $O/train_shards_pool_n$N.log 2>&1; echo "train_shards n$N rc=$?"; grep -E "iter|GPU" $O/train_shards_pool_n$N.log | tail -4
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log; grep -E "passed|failed|^FAILED|^ERROR" $O/pytest_gpu.log | tail -8
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 400 -c 120 --csv --log-file $O/r2_launches_train_shards_pool.csv python tools/train_shards.py --N_rand 1 --hard_ratio 0.2 --hard_mul 1 --steps 60 > $O/train_shards_ncu.log 2>&1; echo "ncu train_shards rc=$?"
ls -la $O
