#!/bin/bash
mkdir -p gpurun_out/r2_11; O=gpurun_out/r2_11
timeout 900 python -m pytest tests -m gpu -q -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
grep -E "passed|failed|rays:|n=|Error|error" $O/pytest_gpu.log | tail -14; tail -3 $O/smoke.log
python - <<PY
import json
d = json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
print("bench value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "kernel_ms", d["roofline"]["kernel_ms"], d["clocks"])
PY
