#!/bin/bash
mkdir -p gpurun_out
{
for E in 1 2 3 4; do echo "== epoch $E"; HPF_HW_EPOCH=$E python profiles/tools/run_solve.py 65536; done
echo "== epoch 2, B=8192"; HPF_HW_EPOCH=2 python profiles/tools/run_solve.py 8192
echo "== epoch 1, B=8192"; HPF_HW_EPOCH=1 python profiles/tools/run_solve.py 8192
echo "== epoch 2, B=262144"; HPF_HW_EPOCH=2 python profiles/tools/run_solve.py 262144
} > gpurun_out/ab_epoch.log 2>&1
cat gpurun_out/ab_epoch.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -k "refill or warp_kernel or nominal or status" > gpurun_out/pytest_ab.log 2>&1
tail -5 gpurun_out/pytest_ab.log
