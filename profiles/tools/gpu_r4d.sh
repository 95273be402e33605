#!/bin/bash
# A/B: w_N on a side stream next to the fundamental stage (default) against w_N after it (HPF_WN_SERIAL=1)
mkdir -p gpurun_out
{
for rep in 1 2; do
for B in 65536 8192; do
  echo "== overlap  B=$B"; python profiles/tools/run_solve.py $B 8 2>&1 | grep -v Warn
  echo "== serial   B=$B"; HPF_WN_SERIAL=1 python profiles/tools/run_solve.py $B 8 2>&1 | grep -v Warn
done
done
echo "== e2e overlap"; python profiles/tools/run_e2e_plans.py 65536 4096,8192,12288,16384 2>&1 | grep -v Warn | head -2
echo "== e2e serial";  HPF_WN_SERIAL=1 python profiles/tools/run_e2e_plans.py 65536 4096,8192,12288,16384 2>&1 | grep -v Warn | head -2
} > gpurun_out/r4_wn_overlap_ab.txt 2>&1
cat gpurun_out/r4_wn_overlap_ab.txt
timeout 600 python -m pytest tests -m gpu -q -x -k "graph or stream or host or guard or refill or nominal" 2>&1 | tail -4
