#!/bin/bash
mkdir -p gpurun_out/r2_14; O=gpurun_out/r2_14
timeout 200 python tools/gpu_chunk_cost.py > $O/chunk_cost.log 2>&1; cat $O/chunk_cost.log | tail -8
