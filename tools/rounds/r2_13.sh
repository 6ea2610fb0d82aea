#!/bin/bash
mkdir -p gpurun_out/r2_13; O=gpurun_out/r2_13
N=${1:-2}
R2L_DDP_SWEEP=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/gpu_ddp_check.py > $O/ddp_check_n$N.log 2>&1; echo "ddp rc=$?"; grep -E "world|Error|error" $O/ddp_check_n$N.log | tail -12
