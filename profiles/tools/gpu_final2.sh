#!/bin/bash
# Bench lines + ncu captures, with the summaries made on the box (the .ncu-rep of 60 kernels with
# sources exceeds the 64 MiB that travel back).  usage: gpu_final2.sh <git-sha>
SHA=${1:-unknown}
mkdir -p gpurun_out /tmp/ncu
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-other > gpurun_out/ncu_launches.log 2>&1; echo "ncu launch list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'harm_hw|fund_tile|wn_lane|mismatch_lane|jacobian_kernel|lu_solve_kernel|solve_kernel|zgemm|wn_tile|harm_cta' \
    -c 60 -f -o /tmp/ncu/r2_kernels python profiles/tools/ncu_targets.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
python profiles/tools/ncu_summary.py /tmp/ncu/r2_kernels.ncu-rep gpurun_out/r2_ncu_kernels.csv $SHA
ls -la /tmp/ncu gpurun_out
