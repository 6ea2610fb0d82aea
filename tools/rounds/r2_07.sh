#!/bin/bash
mkdir -p gpurun_out/r2_07; O=gpurun_out/r2_07
R2L_LIB_OVERRIDE=$PWD/r2l_b200/csrc/libr2l_old.so timeout 120 python tools/gpu_trace.py infer 2 > $O/trace_old.log 2>&1; echo OLD; tail -9 $O/trace_old.log | head -7
timeout 120 python tools/gpu_trace.py infer 2 > $O/trace_new.log 2>&1; echo NEW; tail -9 $O/trace_new.log | head -7
