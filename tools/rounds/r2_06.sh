#!/bin/bash
mkdir -p gpurun_out/r2_06; O=gpurun_out/r2_06
timeout 120 python tools/gpu_trace.py infer 2 > $O/trace_infer.log 2>&1; tail -9 $O/trace_infer.log
timeout 120 python tools/gpu_trace.py train 2 > $O/trace_train.log 2>&1; tail -9 $O/trace_train.log | head -3
timeout 400 python tools/gpu_check_forms.py 1,2 > $O/forms.log 2>&1; echo "forms rc=$?"
timeout 600 python -m pytest tests/test_gpu.py -m gpu -q -s -k "backward or forms or golden" > $O/pytest_sub.log 2>&1; echo "pytest rc=$?"
grep -E "golden|N=4096|N=160000|N=18944" $O/forms.log | tail -12; grep -E "passed|failed|rays:|n=" $O/pytest_sub.log | tail -8
