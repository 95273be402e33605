#!/bin/bash
# compute-sanitizer passes over every persistent / warp-specialised kernel (small ragged batches, grid
# capped at 2 CTAs so that lanes / CTAs refill from the work queue): profiles/tools/sanitize_driver.py
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 700 compute-sanitizer --tool $tool --print-limit 20 python profiles/tools/sanitize_driver.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "exit $?" >> gpurun_out/sanitizer_$tool.txt
done
timeout 600 compute-sanitizer --tool initcheck --print-limit 20 python profiles/tools/sanitize_driver.py warp tile dense standalone > gpurun_out/sanitizer_initcheck.txt 2>&1
echo "exit $?" >> gpurun_out/sanitizer_initcheck.txt
tail -4 gpurun_out/sanitizer_*.txt
