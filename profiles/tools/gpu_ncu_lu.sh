#!/bin/bash
mkdir -p gpurun_out
python profiles/tools/run_lu.py 16384 3 > gpurun_out/lu_plain.log 2>&1 || { cat gpurun_out/lu_plain.log; exit 1; }
HPF_LU_CLASSIC=1 python profiles/tools/run_lu.py 16384 3 >> gpurun_out/lu_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lu_solve_kernel -c 1 -f -o gpurun_out/r2_lu_panel \
    python profiles/tools/run_lu.py 4096 1 > gpurun_out/ncu_lu.log 2>&1
cat gpurun_out/lu_plain.log; tail -n 3 gpurun_out/ncu_lu.log
