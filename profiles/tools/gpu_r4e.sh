#!/bin/bash
# 4-GPU box: topology, host-path probe (D2H alone / together / in pairs), then the 4-GPU bench lines
mkdir -p gpurun_out
{
nvidia-smi topo -m 2>&1 | head -20
lscpu 2>/dev/null | grep -i "numa\|socket\|model name" 
for A in "" 1; do
  echo "== probe, PROBE_AFFINITY=$A"
  PROBE_AFFINITY=$A timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 profiles/tools/d2h_topology.py 2>&1 | grep -v "Warn\|warn\|OMP_NUM\|\*\*\*"
done
} > gpurun_out/r4_host_path_4gpu.txt 2>&1
cat gpurun_out/r4_host_path_4gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu --no-other --no-dense --no-kernels > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; echo "bench 4gpu exit $?"
tail -c 400 gpurun_out/bench_4gpu.err
