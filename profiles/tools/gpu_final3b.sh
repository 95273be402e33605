#!/bin/bash
# Final pass, part B (after part A's ncu summary is committed under profiles/): bench lines + ncu launch list
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-other > gpurun_out/ncu_launches.log 2>&1; echo "ncu launch list exit $?"
python profiles/tools/launch_summary.py gpurun_out/r2_launches.csv > gpurun_out/r2_launch_list_summary.txt
python __graft_entry__.py smoke 2>&1 | tail -1
ls -la gpurun_out
