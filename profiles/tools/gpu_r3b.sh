#!/bin/bash
mkdir -p gpurun_out
HPF_LS_TIMING=1 HPF_LOCKSTEP=1 timeout 300 python profiles/tools/run_other.py radial200 8192 1 2>&1 | grep -v Warn | tee gpurun_out/lockstep_timing2.log
HPF_LS_TIMING=1 HPF_LOCKSTEP=1 timeout 400 python profiles/tools/run_other.py meshed1000 1024 1 2>&1 | grep -v Warn | tee -a gpurun_out/lockstep_timing2.log
