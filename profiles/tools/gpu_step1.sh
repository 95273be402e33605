#!/bin/bash
# GPU box: parity tests, quick A/B timings of the harmonic kernels, compute-sanitizer passes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s --maxfail=12 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
{
for B in 65536 8192; do
  echo "== warp kernel, 2 CTAs/SM, B=$B"; python profiles/tools/run_solve.py $B
  echo "== warp kernel, 1 CTA/SM,  B=$B"; HPF_HW_MINB=1 python profiles/tools/run_solve.py $B
  echo "== tile kernel,            B=$B"; HPF_HARM_KERNEL=tile python profiles/tools/run_solve.py $B
done
echo "== identical scenarios (no individual refills): warp / tile"
HPF_SAME=1 python profiles/tools/run_solve.py 65536
HPF_SAME=1 HPF_HARM_KERNEL=tile python profiles/tools/run_solve.py 65536
} > gpurun_out/ab_timing.log 2>&1
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/tools/sanitize_driver.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "exit $?" >> gpurun_out/sanitizer_$tool.txt
done
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/ab_timing.log; tail -3 gpurun_out/sanitizer_*.txt
