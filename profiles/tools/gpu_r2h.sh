#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -s -k "woodbury or norton_contraction or synthetic_networks or config4 or config5 or guard_bands_large or degenerate" > gpurun_out/pytest_sel6.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_sel6.log
grep -v "^$" gpurun_out/pytest_sel6.log | tail -12
python profiles/tools/run_wn.py radial200 8192 2 2>&1 | grep -v Warn
python profiles/tools/run_wn.py meshed1000 1024 2 2>&1 | grep -v Warn
HPF_SETUP=gj python profiles/tools/run_wn.py meshed1000 1024 2 2>&1 | grep -v Warn
