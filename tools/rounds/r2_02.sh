#!/bin/bash
# Round 2, visit 2: half form with fresh accumulators (Y) + fp32 residual add + small-terms-first order: parity and speed.
mkdir -p gpurun_out/r2_02; O=gpurun_out/r2_02
timeout 900 python -m pytest tests -m gpu -q -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
timeout 120 python tools/gpu_trace.py infer 2 > $O/trace_infer_form2.log 2>&1
timeout 120 python tools/gpu_trace.py train 2 > $O/trace_train_form2.log 2>&1
timeout 400 python tools/gpu_check_forms.py 1,2 > $O/forms.log 2>&1; echo "forms rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
grep -E "passed|failed|rays:|n=" $O/pytest_gpu.log | tail -12; tail -2 $O/smoke.log; tail -8 $O/trace_infer_form2.log; grep -E "golden|N=4096" $O/forms.log | tail -10
python - <<PY
import json
d = json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
print("bench value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "kernel_ms", d["roofline"]["kernel_ms"], d["clocks"])
PY
