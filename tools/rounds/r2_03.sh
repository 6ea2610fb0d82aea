#!/bin/bash
# Round 2, visit 3: debias calibration sweep, new issue order + stream prefetch: speed and accuracy.
mkdir -p gpurun_out/r2_03; O=gpurun_out/r2_03
timeout 600 python tools/gpu_accum_calibrate.py > $O/calibrate.log 2>&1; echo "calibrate rc=$?"
timeout 120 python tools/gpu_trace.py infer 2 > $O/trace_infer_form2.log 2>&1
timeout 120 python tools/gpu_trace.py train 2 > $O/trace_train_form2.log 2>&1
timeout 400 python tools/gpu_check_forms.py 2 > $O/forms.log 2>&1; echo "forms rc=$?"
timeout 600 python -m pytest tests/test_gpu.py -m gpu -q -s -k "backward or forms or golden" > $O/pytest_sub.log 2>&1; echo "pytest rc=$?"
cat $O/calibrate.log; tail -8 $O/trace_infer_form2.log; grep -E "golden|N=4096" $O/forms.log | tail -6; grep -E "passed|failed|rays:|n=" $O/pytest_sub.log | tail -8
