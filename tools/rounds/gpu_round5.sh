#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/gpu_dw_sched.py quick > gpurun_out/dw_quick.log 2>&1; rc=$?; echo "dw quick rc=$rc"; cat gpurun_out/dw_quick.log
if [ $rc -ne 0 ]; then echo "quick check failed: stopping"; exit 1; fi
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python tools/gpu_dw_sched.py > gpurun_out/dw_sched.log 2>&1; echo "dw_sched rc=$?"
timeout 200 python tools/gpu_overlap.py > gpurun_out/overlap.log 2>&1
timeout 300 python tools/gpu_render_bench.py > gpurun_out/render_bench.log 2>&1; echo "render rc=$?"
timeout 300 python bench.py --steps 50 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
grep -E "passed|failed|Error" gpurun_out/pytest_gpu.log | tail -5; cat gpurun_out/dw_sched.log gpurun_out/overlap.log gpurun_out/render_bench.log; cat gpurun_out/bench.json
