#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q > gpurun_out/pytest_train.log 2>&1; rc=$?; echo "pytest train rc=$rc"; tail -40 gpurun_out/pytest_train.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r1_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"r2l_chain_kernel|r2l_dw_kernel" --launch-skip 8 -c 4 -o gpurun_out/r1_full_4096 -f python tools/gpu_profile_target.py 4096 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
