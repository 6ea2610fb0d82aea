#!/bin/bash
# Round 2, visit 1: fp16 planes + loss scale parity, split form (N halves) vs half form: tests, traces, rates, bench.
mkdir -p gpurun_out/r2_01; O=gpurun_out/r2_01
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
timeout 120 python tools/gpu_mma_rate.py > $O/mma_rate.log 2>&1; echo "rate rc=$?"
for f in 2 3; do timeout 120 python tools/gpu_trace.py infer $f > $O/trace_infer_form$f.log 2>&1; done
timeout 120 python tools/gpu_trace.py train 3 > $O/trace_train_form3.log 2>&1
timeout 400 python tools/gpu_check_forms.py 2,3 > $O/forms.log 2>&1; echo "forms rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
R2L_PAIR_MODE=2 timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_form2.json 2> $O/bench_form2.err; echo "bench2 rc=$?"
tail -5 $O/pytest_gpu.log; tail -2 $O/smoke.log; tail -12 $O/trace_infer_form3.log; cat $O/mma_rate.log | tail -6; grep "N=4096" $O/forms.log | tail -8
python - <<PY
import json
for f in ("bench", "bench_form2"):
    try:
        d = json.loads(open("$O/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "kernel_ms", d["roofline"]["kernel_ms"], d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
