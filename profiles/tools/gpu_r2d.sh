#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -x -k "panel_lu or lu_solve or stepwise or nominal or status_words or guard_bands" > gpurun_out/pytest_sel3.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_sel3.log
tail -5 gpurun_out/pytest_sel3.log
python profiles/tools/run_lu.py 16384 3
HPF_LU_CLASSIC=1 python profiles/tools/run_lu.py 16384 3
HPF_SOLVE=dense python profiles/tools/run_solve.py 65536 2
