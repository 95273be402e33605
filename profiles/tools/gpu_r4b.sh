#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/tools/run_e2e_timeline.py 8192 4096,8192,12288,16384 65536 2>&1 | grep -v Warn > gpurun_out/r4_e2e_timeline.txt
cat gpurun_out/r4_e2e_timeline.txt
