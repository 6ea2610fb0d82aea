#!/bin/bash
mkdir -p gpurun_out/r2_20; O=gpurun_out/r2_20
N=${1:-2}
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/gpu_dp_kernel_time.py > $O/dp_kernel_n$N.log 2>&1; echo "rc=$?"; grep -E "^world|rror" $O/dp_kernel_n$N.log | tail -10
