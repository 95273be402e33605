#!/bin/bash
# A/B: CTAs per SM of the per-CTA large-network kernels (threads per CTA x shared-memory budget)
mkdir -p gpurun_out
{
python profiles/tools/run_other.py radial200 2048 2
HPF_GMEM_THREADS=256 HPF_GMEM_SMEM_KB=110 python profiles/tools/run_other.py radial200 2048 2
HPF_GMEM_THREADS=256 python profiles/tools/run_other.py radial200 2048 2
HPF_GMEM_THREADS=128 HPF_GMEM_SMEM_KB=54 python profiles/tools/run_other.py radial200 2048 2
python profiles/tools/run_other.py meshed1000 296 1
HPF_GMEM_THREADS=256 HPF_GMEM_SMEM_KB=110 python profiles/tools/run_other.py meshed1000 296 1
} 2>&1 | grep -v Warn | tee gpurun_out/gmem_ctas_ab.log
