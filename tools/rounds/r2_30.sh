#!/bin/bash
# round-2 evidence visit (1 GPU): full GPU suite, smoke, bench, ncu launch list + full capture, sanitizer logs, pool-step timing
O=gpurun_out/r2_30; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
grep -E "passed|failed|Error|error:|n=|golden 200|abi " $O/pytest_gpu.log | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log; tail -2 $O/smoke.log
timeout 400 python bench.py --steps 50 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -2 $O/bench.err
timeout 300 python tools/train_shards.py --N_rand 1 --hard_ratio 0.2 --hard_mul 2 --steps 400 > $O/train_shards_pool.log 2>&1; echo "train_shards rc=$?"; tail -2 $O/train_shards_pool.log
timeout 200 python tools/gpu_teacher_frame.py > $O/teacher_frame.log 2>&1; echo "teacher frame rc=$?"; tail -2 $O/teacher_frame.log
timeout 200 python tools/gpu_small_kernels.py > $O/small_kernels.log 2>&1; echo "small kernels rc=$?"; tail -6 $O/small_kernels.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r2_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 > $O/bench_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"r2l_chain_kernel|r2l_dw_kernel" --launch-skip 8 -c 4 -o $O/r2_full_4096 -f python tools/gpu_profile_target.py 4096 > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 $O/ncu_full.log
timeout 120 ncu -i $O/r2_full_4096.ncu-rep --page raw --csv > $O/r2_full_4096_raw.csv 2> $O/ncu_export.err; echo "ncu export rc=$?"
timeout 200 ncu --set full --clock-control none -k regex:"raw2outputs|teacher_kernel|pool_update|pool_draw|sample_pdf" -c 6 -o $O/r2_full_small -f python tools/gpu_small_kernels.py ncu > $O/ncu_small.log 2>&1; echo "ncu small rc=$?"
timeout 120 ncu -i $O/r2_full_small.ncu-rep --page raw --csv > $O/r2_full_small_raw.csv 2>> $O/ncu_export.err
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/gpu_profile_target.py 512 1 > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/sanitizer_memcheck.log
timeout 240 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/gpu_profile_target.py 512 1 > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 $O/sanitizer_racecheck.log
python - <<PY
import json
try:
    d = json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
    print("bench value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"].get("frac"), d["roofline"].get("kernel_ms"), d.get("gpu_launches"), d["clocks"])
except Exception as e:
    print("bench parse failed", e)
PY
rm -f $O/r2_full_small.ncu-rep
ls -la $O
