#!/bin/bash
# Final pass, part A: full GPU test suite + ncu captures (summaries made on the box).  usage: gpu_final3a.sh <git-sha>
SHA=${1:-unknown}
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'harm_hw|fund_tile|wn_lane|mismatch_lane|jacobian_kernel|lu_solve_kernel|solve_kernel|zgemm|wn_tile|harm_cta' \
    -c 60 -f -o /tmp/ncu/r2_kernels python profiles/tools/ncu_targets.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
python profiles/tools/ncu_summary.py /tmp/ncu/r2_kernels.ncu-rep gpurun_out/r2_ncu_kernels.csv $SHA
timeout 600 ncu --set full --clock-control none -k regex:'ls_' -c 10 -f -o /tmp/ncu/r2_ls_lu \
    python profiles/tools/run_lu_batched.py 512 1 > gpurun_out/ncu_ls_lu.log 2>&1; echo "ncu lock-step LU exit $?"
python profiles/tools/ncu_summary.py /tmp/ncu/r2_ls_lu.ncu-rep gpurun_out/r2_ncu_lockstep_kernels.csv $SHA
timeout 600 ncu --set full --clock-control none -k regex:'ls_mismatch|ls_border|ls_uf|ls_pack|ls_compact|ls_backsub|ls_fund' -c 9 -f -o /tmp/ncu/r2_ls_rounds \
    python profiles/tools/run_other.py radial200 1024 1 > gpurun_out/ncu_ls_rounds.log 2>&1; echo "ncu lock-step rounds exit $?"
python profiles/tools/ncu_summary.py /tmp/ncu/r2_ls_rounds.ncu-rep gpurun_out/r2_ncu_lockstep_rounds.csv $SHA
HPF_LS_TIMING=1 python profiles/tools/run_other.py radial200 8192 1 2>&1 | grep -v Warn | tail -3 > gpurun_out/r2_lockstep_phase_timing.txt
HPF_LS_TIMING=1 python profiles/tools/run_other.py meshed1000 1024 1 2>&1 | grep -v Warn | tail -3 >> gpurun_out/r2_lockstep_phase_timing.txt
cat gpurun_out/r2_lockstep_phase_timing.txt
