#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "batched_lu or lockstep" > gpurun_out/pytest_lockstep.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_lockstep.log
tail -15 gpurun_out/pytest_lockstep.log
HPF_LS_TIMING=1 timeout 300 python profiles/tools/run_other.py radial200 8192 1 2>&1 | grep -v Warn | tee gpurun_out/lockstep_timing3.log
HPF_LS_TIMING=1 timeout 400 python profiles/tools/run_other.py meshed1000 1024 1 2>&1 | grep -v Warn | tee -a gpurun_out/lockstep_timing3.log
HPF_LS_NO_PAIR=1 HPF_LS_TIMING=1 timeout 400 python profiles/tools/run_other.py meshed1000 1024 1 2>&1 | grep -v Warn | tee -a gpurun_out/lockstep_timing3.log
