#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for lib in profiles/tools/_bin/libhpf_old.so harmonic_power_flow_b200/libhpf_b200.so; do
  echo "== $lib"
  HPF_LIB=$PWD/$lib HPF_LS_TIMING=1 timeout 300 python profiles/tools/run_other.py radial200 8192 1 2>&1 | grep -v Warn | tail -2 | cut -c1-330
done
done
for lib in profiles/tools/_bin/libhpf_old.so harmonic_power_flow_b200/libhpf_b200.so; do
  echo "== $lib"
  HPF_LIB=$PWD/$lib HPF_LS_TIMING=1 timeout 300 python profiles/tools/run_other.py meshed1000 1024 1 2>&1 | grep -v Warn | tail -2 | cut -c1-330
done
