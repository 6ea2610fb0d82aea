#!/bin/bash
# One GPU-box visit = a list of steps; every step writes into gpurun_out/<tag>/ and never aborts the visit.
#   gpurun --timeout 700 -- 'bash tools/gpu_visit.sh r2_xx tests smoke bench launches ncu_chain ncu_small sanitizer'
#   gpurun --gpus 2 --timeout 480 -- 'bash tools/gpu_visit.sh r2_xx sweep:2 ddp:2 train_pool:2'
# The files that are evidence get copied to profiles/ by hand (profiles/r2_summary.md lists them).
TAG=$1; shift
O=gpurun_out/$TAG; mkdir -p $O
trun() { local n=$1 port=$2; shift 2; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@"; }
for step in "$@"; do
  name=${step%%:*}; N=${step#*:}; [ "$N" = "$step" ] && N=1
  case $name in
    tests)      timeout 900 python -m pytest tests -m gpu -q -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
                grep -E "passed|failed|^FAILED|^ERROR|n=|golden 200|abi " $O/pytest_gpu.log | tail -24 ;;
    smoke)      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log; tail -1 $O/smoke.log ;;
    bench)      if [ $N = 1 ]; then timeout 400 python bench.py --steps 50 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
                else timeout 400 bash -c "$(declare -f trun); trun $N 29532 bench.py --gpus $N --steps 50 --warmup 3" > $O/bench_n$N.json 2> $O/bench_n$N.err; fi
                echo "bench n$N rc=$?"; tail -c 400 $O/bench_n$N.json ;;
    launches)   timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 > $O/bench_ncu.log 2>&1; echo "ncu launches rc=$?" ;;
    ncu_chain)  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"r2l_chain_kernel|r2l_dw_kernel" --launch-skip 8 -c 4 -o $O/full_4096 -f python tools/gpu_profile_target.py 4096 > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
                timeout 120 ncu -i $O/full_4096.ncu-rep --page raw --csv > $O/full_4096_raw.csv 2> $O/ncu_export.err ;;
    ncu_small)  timeout 200 ncu --set full --clock-control none -k regex:"raw2outputs|teacher_kernel|pool_update|pool_draw|sample_pdf" -c 16 -o $O/full_small -f python tools/gpu_small_kernels.py ncu > $O/ncu_small.log 2>&1; echo "ncu small rc=$?"
                timeout 120 ncu -i $O/full_small.ncu-rep --page raw --csv > $O/full_small_raw.csv 2>> $O/ncu_export.err; rm -f $O/full_small.ncu-rep ;;
    small)      timeout 200 python tools/gpu_small_kernels.py > $O/small_kernels.log 2>&1; echo "small kernels rc=$?"; tail -8 $O/small_kernels.log ;;
    teacher)    timeout 200 python tools/gpu_teacher_frame.py > $O/teacher_frame.log 2>&1; echo "teacher frame rc=$?"; tail -1 $O/teacher_frame.log ;;
    sanitizer)  timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/gpu_profile_target.py 512 1 > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
                timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"
                timeout 240 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/gpu_profile_target.py 512 1 > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 $O/sanitizer_racecheck.log ;;
    sweep)      if [ $N = 1 ]; then timeout 300 python tools/render_sweep.py > $O/render_sweep_n1.log 2>&1
                else timeout 300 bash -c "$(declare -f trun); trun $N 29541 tools/render_sweep.py" > $O/render_sweep_n$N.log 2>&1; fi
                echo "render sweep n$N rc=$?"; grep -E "^\{" $O/render_sweep_n$N.log | cut -c1-250 ;;
    ddp)        R2L_DDP_SWEEP=${R2L_DDP_SWEEP:-0} timeout 300 bash -c "$(declare -f trun); trun $N 29531 tools/gpu_ddp_check.py" > $O/ddp_check_n$N.log 2>&1; echo "ddp n$N rc=$?"; grep -E "world|Error|error" $O/ddp_check_n$N.log | tail -8 ;;
    train_pool) if [ $N = 1 ]; then timeout 300 python tools/train_shards.py --N_rand 1 --hard_ratio 0.2 --hard_mul 2 --steps 400 > $O/train_shards_pool_n1.log 2>&1
                else timeout 300 bash -c "$(declare -f trun); trun $N 29542 tools/train_shards.py --N_rand 1 --hard_ratio 0.2 --hard_mul 2 --steps 300" > $O/train_shards_pool_n$N.log 2>&1; fi
                echo "train_shards n$N rc=$?"; grep -E "iter|GPU" $O/train_shards_pool_n$N.log | tail -3 ;;
    launches_pool) timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 400 -c 120 --csv --log-file $O/launches_train_shards_pool.csv python tools/train_shards.py --N_rand 1 --hard_ratio 0.2 --hard_mul 1 --steps 60 > $O/train_shards_ncu.log 2>&1; echo "ncu train_shards rc=$?" ;;
    *) echo "unknown step $step" ;;
  esac
done
ls -la $O
