#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -x -q > gpurun_out/pytest_train.log 2>&1; rc=$?; echo "pytest train rc=$rc"; tail -30 gpurun_out/pytest_train.log
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_train.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 50 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -5 gpurun_out/bench.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json
