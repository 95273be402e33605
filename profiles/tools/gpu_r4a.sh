#!/bin/bash
# chunk-plan sweep of hpf_solve_host (one process), then the host-call tests
mkdir -p gpurun_out
timeout 400 python profiles/tools/run_e2e_plans.py 65536 2>&1 | grep -v Warn > gpurun_out/r4_e2e_plans.txt
tail -8 gpurun_out/r4_e2e_plans.txt
timeout 300 python -m pytest tests -m gpu -q -x -k "host" 2>&1 | tail -3
