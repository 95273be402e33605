#!/bin/bash
# GPU box: parity tests, quick A/B timings of the harmonic kernels, the default bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s --maxfail=12 --durations=15 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
{
for B in 65536 8192; do
  echo "== warp kernel, B=$B"; python profiles/tools/run_solve.py $B
  echo "== tile kernel, B=$B"; HPF_HARM_KERNEL=tile python profiles/tools/run_solve.py $B
done
} > gpurun_out/ab_timing.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench exit $?"
tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/ab_timing.log; tail -c 3000 gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
