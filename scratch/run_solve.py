import sys, os, time, tempfile
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "oracle"))
import numpy as np, torch
import helpers
from harmonic_power_flow_b200 import BatchSolver, scenarios
net, st, _ = helpers.packed_from_files("net3", 25, True, tempfile.mkdtemp())
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
sol = BatchSolver(net)
P, Q, I_N = scenarios.make_batch(net, B, "tight", exact_prefix=64)
dP, dQ, dI = sol.prepare(P, Q, I_N)
r = sol.solve(dP, dQ, dI); torch.cuda.synchronize()
sol.set_profiling(True)
best = 1e9
for _ in range(5):
    r = sol.solve(dP, dQ, dI, out=r); k = sol.last_kernel_ms(); best = min(best, k[1])
print("B=%d harm stage %.3f ms  (%.1f M solves/s)  mean it %.2f checksum %.12e" % (B, best, B / best / 1e3, r.n_iter_h.double().mean().item(), float(r.V_m.sum())))
