"""The Norton contraction w_N = W_NL I_N of a whole batch (hpf_norton_wn) on a BASELINE configuration -
A/B of the CUDA-core kernel (wn_tile_kernel, $HPF_WN_KERNEL=fma) against the FP64 tensor-core GEMM
(zgemm_dmma_kernel, $HPF_WN_KERNEL=dmma).  usage: run_wn.py radial200|meshed1000|net1 [B] [reps]"""
import os
import sys
import time

R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (R, os.path.join(R, "tests"), os.path.join(R, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import torch
import bench
from harmonic_power_flow_b200 import BatchSolver, scenarios

kind = sys.argv[1] if len(sys.argv) > 1 else "radial200"
B = int(sys.argv[2]) if len(sys.argv) > 2 else (8192 if kind == "radial200" else 1024)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
net = bench.load_other(kind)
sol = BatchSolver(net)
t0 = time.perf_counter(); info = sol.struct_info(); torch.cuda.synchronize(); t_setup = time.perf_counter() - t0
P, Q, I_N = scenarios.make_batch(net, B, "tight")
dI = sol._dev(I_N, torch.complex128)
w = sol.norton_wn(dI)
torch.cuda.synchronize()
best = 1e9
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); w = sol.norton_wn(dI); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
nZ, qH = info["nZ"], net.q * net.H
fl = 8.0 * nZ * qH * B
print("%s kernel=%s: nZ=%d qH=%d B=%d  w_N %.3f ms (incl. %.3f ms copy-out)  %.2f TFLOP/s  setup %.2f s  checksum %.12e" % (
    kind, os.environ.get("HPF_WN_KERNEL", "default"), nZ, qH, B, best, 16.0 * nZ * B / 5e9 * 1e3 * 2 / 1e3, fl / best / 1e9,
    t_setup, float(w.real.sum() + w.imag.sum())))
