#!/bin/bash
# Final GPU pass of a round: full GPU test-suite, smoke(), the default bench line, the ncu launch list of a
# short bench run and ONE ncu --set full capture of every kernel of the roofline discussion.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -14 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-other > gpurun_out/ncu_launches.log 2>&1; echo "ncu launch list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'harm_hw|fund_tile|wn_lane|mismatch_lane|jacobian_kernel|lu_solve_kernel|solve_kernel|zgemm|wn_tile|harm_cta' \
    -c 60 -f -o gpurun_out/r2_kernels python profiles/tools/ncu_targets.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
tail -3 gpurun_out/ncu_full.log
