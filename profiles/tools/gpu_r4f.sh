#!/bin/bash
# 2-GPU bench lines of the final binary (weak scaling + the strong-scaling split of the north-star target)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-other --no-dense --no-kernels > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench 2gpu exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 10 --warmup 3 --strong --no-cpu --no-other --no-dense --no-kernels > gpurun_out/bench_2gpu_strong.json 2> gpurun_out/bench_2gpu_strong.err; echo "bench 2gpu strong exit $?"
tail -c 300 gpurun_out/bench_2gpu.err
