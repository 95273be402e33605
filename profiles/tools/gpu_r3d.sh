#!/bin/bash
mkdir -p gpurun_out
python profiles/tools/run_lu_batched.py 512 2 2>&1 | grep -v Warn | tee gpurun_out/lu_batched.log
HPF_LS_NO_PAIR=1 python profiles/tools/run_lu_batched.py 512 2 2>&1 | grep -v Warn | tee -a gpurun_out/lu_batched.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ls_' -c 9 -f -o gpurun_out/r3_ls_lu \
    python profiles/tools/run_lu_batched.py 512 1 > gpurun_out/ncu_ls_lu.log 2>&1
tail -n 3 gpurun_out/ncu_ls_lu.log; ls -la gpurun_out
