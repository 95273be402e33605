"""ONE launch of every kernel the roofline discussion refers to, for a single `ncu --set full` capture:

    ncu --set full --clock-control none --import-source on \
        -k regex:'harm_hw|fund_tile|wn_lane|mismatch_lane|jacobian_kernel|lu_solve_kernel|solve_kernel|zgemm|wn_tile|harm_cta' \
        -c 60 -f -o gpurun_out/r2_kernels python profiles/tools/ncu_targets.py

Config 3 (net3, 65,536 scenarios): fund_tile_kernel, wn_lane_kernel, harm_hw_kernel (hpf_solve),
mismatch_lane_kernel, jacobian_kernel (16,384), lu_solve_kernel<0> (4,096, panel LU), solve_kernel<0>
(4,096, fused dense solve).  Config 4 network (200-bus feeder, 296 scenarios = 2 per CTA):
solve_kernel<1|2> (fundamental stage), zgemm_dmma_kernel (w_N), harm_cta_kernel; wn_tile_kernel on
32 scenarios (the CUDA-core Norton contraction)."""
import os
import sys
import tempfile

R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (R, os.path.join(R, "tests"), os.path.join(R, "oracle")):
    sys.path.insert(0, p)
import torch
import bench
import helpers
from harmonic_power_flow_b200 import BatchSolver, scenarios

net, _, _ = helpers.packed_from_files("net3", 25, True, tempfile.mkdtemp())
sol = BatchSolver(net)
B = 65536
P, Q, I_N = scenarios.make_batch(net, B, "tight")
dP, dQ, dI = sol.prepare(P, Q, I_N)
r = sol.solve(dP, dQ, dI)
raw = sol.solve(dP[:, :16384].contiguous(), dQ[:, :16384].contiguous(), dI[:, :, :16384].contiguous(), raw=True,
                max_iter_h=3, want_I_inj=False)
f, err = sol.mismatch(r.V_m, r.V_a, dP, dQ, dI)
J = sol.jacobian(raw.V_m, raw.V_a)
f16, _ = sol.mismatch(raw.V_m, raw.V_a, dP[:, :16384].contiguous(), dQ[:, :16384].contiguous(), dI[:, :, :16384].contiguous())
dx, info = sol.lu_solve(J[:4096].contiguous(), f16[:, :4096].contiguous())
d = sol.solve(dP[:, :4096].contiguous(), dQ[:, :4096].contiguous(), dI[:, :, :4096].contiguous(), dense=True)
torch.cuda.synchronize()
sol.close()

os.environ["HPF_GJ_UNBLOCKED"] = "1"          # (keeps the set-up's panel GEMMs out of the capture)
os.environ["HPF_LOCKSTEP"] = "0"              # (per-CTA kernels; the lock-step kernels: run_lu_batched.py / run_other.py)
net4 = bench.load_other("radial200")
sol4 = BatchSolver(net4)
P4, Q4, I4 = scenarios.make_batch(net4, 296, "tight")
r4 = sol4.solve(P4, Q4, I4)
w = sol4.norton_wn(I4[:, :, :32].copy())       # B < 64: wn_tile_kernel
torch.cuda.synchronize()
print("converged", int((r.status == 0).sum()), int((d.status == 0).sum()), int((r4.status == 0).sum()))
