#!/bin/bash
# one ncu --set full capture of the harmonic Newton kernel (config 3, 65,536 scenarios), source-level,
# plus one of the 8,192-scenario launch (the 8-GPU strong-scaling share)
mkdir -p gpurun_out
python profiles/tools/run_solve.py 65536 2 > gpurun_out/ncu_harm_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:harm_hw_kernel -s 1 -c 1 -f -o gpurun_out/r2_harm_hw \
    python profiles/tools/run_solve.py 65536 1 > gpurun_out/ncu_harm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:harm_hw_kernel -s 1 -c 1 -f -o gpurun_out/r2_harm_hw_8192 \
    python profiles/tools/run_solve.py 8192 1 > gpurun_out/ncu_harm_8192.log 2>&1
tail -3 gpurun_out/ncu_harm.log gpurun_out/ncu_harm_8192.log
