#!/bin/bash
mkdir -p gpurun_out/r2_09; O=gpurun_out/r2_09
timeout 120 python tools/gpu_trace.py infer 2 > $O/trace_infer.log 2>&1; tail -16 $O/trace_infer.log
