#!/bin/bash
mkdir -p gpurun_out
for B in 65536 32768 16384 8192; do
  echo "==== B = $B"
  timeout 400 python profiles/tools/run_e2e_plans.py $B 2>&1 | grep -v Warn
done > gpurun_out/r4_e2e_plans2.txt
grep -A4 "====\|best by" gpurun_out/r4_e2e_plans2.txt
timeout 300 python profiles/tools/run_e2e_timeline.py 8192 4096,8192,12288,16384 2>&1 | grep -v Warn > gpurun_out/r4_e2e_timeline2.txt
cat gpurun_out/r4_e2e_timeline2.txt
timeout 300 python -m pytest tests -m gpu -q -x -k "host" 2>&1 | tail -3
