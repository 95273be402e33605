#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "norton_contraction or reference_shaped or config5 or synthetic_networks" > gpurun_out/pytest_sel5.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_sel5.log
tail -5 gpurun_out/pytest_sel5.log
for K in fma dmma; do
  HPF_WN_KERNEL=$K python profiles/tools/run_wn.py radial200 8192 3
  HPF_WN_KERNEL=$K python profiles/tools/run_wn.py meshed1000 1024 3
  HPF_WN_KERNEL=$K python profiles/tools/run_wn.py net1 4096 3
done 2>&1 | grep -v Warning | tee gpurun_out/wn_ab.log
