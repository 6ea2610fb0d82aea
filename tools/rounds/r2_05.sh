#!/bin/bash
mkdir -p gpurun_out/r2_05; O=gpurun_out/r2_05
for v in 0 1 2 3 4 5 6 7; do timeout 120 python tools/gpu_trace.py infer 2 $v > $O/trace_v$v.log 2>&1; echo "variant $v"; tail -9 $O/trace_v$v.log | head -6; done
