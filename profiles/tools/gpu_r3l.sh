#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "batched_lu or lockstep" > gpurun_out/pytest_lockstep.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_lockstep.log
tail -2 gpurun_out/pytest_lockstep.log
python profiles/tools/run_lu_batched.py 512 2 2>&1 | grep -v Warn | tee gpurun_out/lu_batched.log
HPF_LS_TIMING=1 timeout 300 python profiles/tools/run_other.py radial200 8192 1 2>&1 | grep -v Warn | tail -3 | tee gpurun_out/lockstep_timing7.log
HPF_LS_TIMING=1 timeout 400 python profiles/tools/run_other.py meshed1000 1024 1 2>&1 | grep -v Warn | tail -3 | tee -a gpurun_out/lockstep_timing7.log
