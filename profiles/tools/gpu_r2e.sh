#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "panel_lu or lu_solve or stepwise" > gpurun_out/pytest_sel4.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_sel4.log
tail -4 gpurun_out/pytest_sel4.log
python profiles/tools/run_lu.py 16384 3
