#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ls_panel -s 20 -c 1 -f -o gpurun_out/r3_ls_panel1 \
    python profiles/tools/run_lu_batched.py 512 1 > gpurun_out/ncu_ls_panel1.log 2>&1
tail -n 2 gpurun_out/ncu_ls_panel1.log
