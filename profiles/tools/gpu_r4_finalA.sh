#!/bin/bash
# Final pass of the last session, part A: full GPU test suite + ncu --set full summary of the headline kernels
mkdir -p gpurun_out /tmp/ncu
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'harm_hw|fund_tile|wn_lane|mismatch_lane|jacobian_kernel|lu_solve_kernel|solve_kernel|zgemm|wn_tile|harm_cta' \
    -c 60 -f -o /tmp/ncu/r2_kernels python profiles/tools/ncu_targets.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
python profiles/tools/ncu_summary.py /tmp/ncu/r2_kernels.ncu-rep gpurun_out/r2_ncu_kernels.csv c7b0551
head -c 300 gpurun_out/r2_ncu_kernels.csv; wc -l gpurun_out/r2_ncu_kernels.csv
