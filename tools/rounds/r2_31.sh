#!/bin/bash
# round-2 verification visit (1 GPU): the whole GPU suite, smoke, config-5 sweep on one GPU, memcheck over smoke(), ncu pages of the small kernels
O=gpurun_out/r2_31; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
grep -E "passed|failed|^FAILED|^ERROR|n=|golden 200|abi " $O/pytest_gpu.log | tail -24
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log; tail -1 $O/smoke.log
timeout 300 python tools/render_sweep.py > $O/render_sweep_n1.log 2>&1; echo "render sweep rc=$?"; grep -E "^\{" $O/render_sweep_n1.log | cut -c1-260
timeout 200 python tools/gpu_small_kernels.py > $O/small_kernels.log 2>&1; echo "small kernels rc=$?"; tail -7 $O/small_kernels.log
timeout 200 ncu --set full --clock-control none -k regex:"raw2outputs|teacher_kernel|pool_update|pool_draw|sample_pdf" -c 16 -o $O/r2_full_small -f python tools/gpu_small_kernels.py ncu > $O/ncu_small.log 2>&1; echo "ncu small rc=$?"
timeout 120 ncu -i $O/r2_full_small.ncu-rep --page raw --csv > $O/r2_full_small_raw.csv 2> $O/ncu_export.err
rm -f $O/r2_full_small.ncu-rep
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -3 $O/sanitizer_memcheck_smoke.log
ls -la $O
