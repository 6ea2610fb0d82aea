#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python tools/gpu_dw_sched.py > gpurun_out/dw_sched.log 2>&1; echo "dw_sched rc=$?"
for m in infer train; do timeout 120 python tools/gpu_trace.py $m 2 > gpurun_out/trace_${m}_form2.log 2>&1; done
timeout 120 python tools/gpu_trace.py infer 0 > gpurun_out/trace_infer_form0.log 2>&1
timeout 200 python tools/gpu_stats.py > gpurun_out/stats.log 2>&1
timeout 200 python tools/gpu_overlap.py > gpurun_out/overlap.log 2>&1
timeout 300 python bench.py --steps 50 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/dw_sched.log; cat gpurun_out/bench.json
