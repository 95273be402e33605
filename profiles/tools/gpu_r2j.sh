#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python profiles/tools/run_hist.py 2>&1 | grep -v Warn | tee gpurun_out/iter_hist.log
python profiles/tools/run_other.py radial200 2048 2 2>&1 | grep -v Warn | tee -a gpurun_out/gmem_ctas_ab.log
