#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "batched_lu or lockstep" > gpurun_out/pytest_lockstep.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_lockstep.log
tail -3 gpurun_out/pytest_lockstep.log
HPF_LS_TIMING=1 timeout 300 python profiles/tools/run_other.py radial200 8192 1 2>&1 | grep -v Warn | tail -3 | tee gpurun_out/lockstep_timing6.log
HPF_LS_TIMING=1 timeout 400 python profiles/tools/run_other.py meshed1000 1024 1 2>&1 | grep -v Warn | tail -3 | tee -a gpurun_out/lockstep_timing6.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ls_panel' -c 2 -f -o gpurun_out/r3_ls_panel \
    python profiles/tools/run_lu_batched.py 512 1 > gpurun_out/ncu_ls_panel.log 2>&1
tail -n 2 gpurun_out/ncu_ls_panel.log
