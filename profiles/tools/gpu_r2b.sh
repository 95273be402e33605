#!/bin/bash
# GPU box: previously failing tests, small-batch A/B of the harmonic kernel, the default bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -x -k "large_networks_structured or full_size_batch or config4 or host_entry or nominal" > gpurun_out/pytest_sel.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_sel.log
{
for B in 8192 16384 32768; do
  echo "== default B=$B"; python profiles/tools/run_solve.py $B
  echo "== MINB=1 B=$B"; HPF_HW_MINB=1 python profiles/tools/run_solve.py $B
  echo "== MAX_CTAS=148 B=$B"; HPF_MAX_CTAS=148 python profiles/tools/run_solve.py $B
done
echo "== MINB=1 B=65536"; HPF_HW_MINB=1 python profiles/tools/run_solve.py 65536
} > gpurun_out/ab_small.log 2>&1
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench exit $?"
tail -8 gpurun_out/pytest_sel.log; cat gpurun_out/ab_small.log; tail -5 gpurun_out/bench_default.err
