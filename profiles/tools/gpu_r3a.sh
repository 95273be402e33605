#!/bin/bash
# lock-step path: its tests first, then timing of configs 4 / 5 with the path on and off
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "batched_lu or lockstep" > gpurun_out/pytest_lockstep.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_lockstep.log
tail -25 gpurun_out/pytest_lockstep.log
HPF_LOCKSTEP=1 timeout 300 python profiles/tools/run_other.py radial200 8192 1 2>&1 | grep -v Warn | tee gpurun_out/lockstep_timing.log
HPF_LOCKSTEP=0 timeout 300 python profiles/tools/run_other.py radial200 8192 1 2>&1 | grep -v Warn | tee -a gpurun_out/lockstep_timing.log
HPF_LOCKSTEP=1 timeout 400 python profiles/tools/run_other.py meshed1000 1024 1 2>&1 | grep -v Warn | tee -a gpurun_out/lockstep_timing.log
