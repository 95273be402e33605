import sys, os, time, tempfile
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "oracle"))
import numpy as np, torch
import helpers
from harmonic_power_flow_b200 import BatchSolver, scenarios
net, st, _ = helpers.packed_from_files("net3", 25, True, tempfile.mkdtemp())
B = 16384
sol = BatchSolver(net)
P, Q, I_N = scenarios.make_batch(net, B, "tight", exact_prefix=64)
dP, dQ, dI = sol.prepare(P, Q, I_N)
raw = sol.solve(dP, dQ, dI, raw=True, max_iter_h=3, want_I_inj=False)
J = sol.jacobian(raw.V_m, raw.V_a)
f, _ = sol.mismatch(raw.V_m, raw.V_a, dP, dQ, dI)
dx, info = sol.lu_solve(J, f); torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dx, info = sol.lu_solve(J, f); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
N = sol.N
fl = 2.0 / 3.0 * N ** 3 + 2.0 * N * N
# residual check on a few scenarios
Jv = sol.jacobian_view(J)
res = (torch.einsum("bij,jb->ib", Jv[:64], dx[:, :64]) - f[:, :64]).abs().max().item()
print("lu_solve N=%d B=%d: %.3f ms  %.2f TFLOP/s  info!=0: %d  max residual (64 scen) %.2e  |dx| max %.2e" % (
    N, B, best, fl * B / best / 1e9, int((info != 0).sum()), res, dx.abs().max().item()))
r = sol.solve(dP, dQ, dI, dense=True); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); r = sol.solve(dP, dQ, dI, dense=True, out=r); e1.record(); torch.cuda.synchronize()
print("dense fused solve B=%d: %.1f ms  %.0f solves/s  conv %d  mean it %.2f" % (B, e0.elapsed_time(e1), B / e0.elapsed_time(e1) * 1e3, int((r.status == 0).sum()), r.n_iter_h.double().mean().item()))
