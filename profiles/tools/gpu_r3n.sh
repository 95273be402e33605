#!/bin/bash
mkdir -p gpurun_out
for cfg in "" "HPF_HW_MINB=1" "HPF_MAX_CTAS=148" "HPF_MAX_CTAS=200" "HPF_HW_MINB=1 HPF_HW_EPOCH=1" "HPF_MAX_CTAS=148 HPF_HW_EPOCH=1" "HPF_HW_EPOCH=1"; do
  echo "== $cfg" | tee -a gpurun_out/small_batch_ab.log
  env $cfg python profiles/tools/run_solve.py 8192 20 2>&1 | grep -v Warn | tail -1 | tee -a gpurun_out/small_batch_ab.log
done
