#!/bin/bash
# quick A/B of the harmonic kernels + the parity tests that touch them
mkdir -p gpurun_out
{
for B in 65536 8192; do
  echo "== warp kernel, 2 CTAs/SM, B=$B"; python profiles/tools/run_solve.py $B
  echo "== warp kernel, 1 CTA/SM,  B=$B"; HPF_HW_MINB=1 python profiles/tools/run_solve.py $B
done
echo "== tile kernel B=65536"; HPF_HARM_KERNEL=tile python profiles/tools/run_solve.py 65536
echo "== identical scenarios: warp"; HPF_SAME=1 python profiles/tools/run_solve.py 65536
} > gpurun_out/ab_timing.log 2>&1
cat gpurun_out/ab_timing.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -k "refill or warp_kernel" > gpurun_out/pytest_ab.log 2>&1
tail -15 gpurun_out/pytest_ab.log
