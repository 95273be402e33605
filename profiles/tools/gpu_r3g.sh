#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/tools/run_lockstep_rates.py meshed 70 150 2>&1 | grep -v Warn | tail -3 | tee gpurun_out/lockstep_rates.log
timeout 600 python profiles/tools/run_lockstep_rates.py radial 40 150 2>&1 | grep -v Warn | tail -3 | tee -a gpurun_out/lockstep_rates.log
