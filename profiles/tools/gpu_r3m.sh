#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "degenerate or batched_lu or lockstep_waves or guard_bands_large" 2>&1 | tail -4
HPF_LS_TIMING=1 timeout 300 python profiles/tools/run_other.py radial200 8192 1 2>&1 | grep -v Warn | tail -2
