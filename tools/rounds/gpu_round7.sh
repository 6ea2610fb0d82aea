#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|assert" gpurun_out/pytest_gpu.log | tail -8
timeout 300 python tools/gpu_teacher_frame.py > gpurun_out/teacher_frame.log 2>&1; echo "teacher rc=$?"; tail -3 gpurun_out/teacher_frame.log
timeout 300 python tools/train_shards.py --N_rand 1 --steps 300 > gpurun_out/train_shards_1.log 2>&1; echo "train1 rc=$?"; tail -4 gpurun_out/train_shards_1.log
timeout 300 python tools/train_shards.py --N_rand 20 --steps 60 --hard_mul 2 > gpurun_out/train_shards_20.log 2>&1; echo "train20 rc=$?"; tail -3 gpurun_out/train_shards_20.log
