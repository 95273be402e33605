#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ls_update_kernel' -s 1 -c 1 -f -o gpurun_out/r3_ls_update \
    python profiles/tools/run_lu_batched.py 512 1 > gpurun_out/ncu_ls_update.log 2>&1
tail -n 2 gpurun_out/ncu_ls_update.log
