#!/bin/bash
# Final pass, part B (part A's ncu summary is in profiles/): bench lines + ncu launch list + smoke
mkdir -p gpurun_out
( time timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2> gpurun_out/bench_default.time; echo "bench exit $?"; tail -3 gpurun_out/bench_default.time
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2> gpurun_out/bench_reference.time; echo "bench reference exit $?"; tail -3 gpurun_out/bench_reference.time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-other > gpurun_out/ncu_launches.log 2>&1; echo "ncu launch list exit $?"
python profiles/tools/launch_summary.py gpurun_out/r2_launches.csv > gpurun_out/r2_launch_list_summary.txt
python __graft_entry__.py smoke 2>&1 | tail -1
ls -la gpurun_out
