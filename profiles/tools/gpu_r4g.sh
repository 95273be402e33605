#!/bin/bash
# 8-GPU box: topology + host-path probe only
mkdir -p gpurun_out
{
nvidia-smi topo -m 2>&1 | head -12 | cut -c1-150
lscpu 2>/dev/null | grep -i "numa\|socket\|model name"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 profiles/tools/d2h_topology.py 2>&1 | grep -v "Warn\|warn\|OMP_NUM\|\*\*\*"
} > gpurun_out/r4_host_path_8gpu.txt 2>&1
cat gpurun_out/r4_host_path_8gpu.txt
