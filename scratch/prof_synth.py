import sys, os, time, tempfile
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "oracle"))
import numpy as np, torch
import helpers
from harmonic_power_flow_b200 import BatchSolver, scenarios
kind = sys.argv[1]; n = int(sys.argv[2]); scale = float(sys.argv[3]); hmax = int(sys.argv[4]); B = int(sys.argv[5])
net, st = helpers.synthetic_packed(kind, tempfile.mkdtemp(), h_max=hmax, n=n, load_scale=scale)
sol = BatchSolver(net)
P, Q, I_N = scenarios.make_batch(net, B, "tight", exact_prefix=8)
dP, dQ, dI = sol.prepare(P, Q, I_N)
sol.set_profiling(True) if hasattr(sol, "set_profiling") else None
for _ in range(2):
    t = time.time(); r = sol.solve(dP, dQ, dI); torch.cuda.synchronize(); dt = time.time() - t
print("B=%d solve %.3fs %.1f solves/s mean it %.1f" % (B, dt, B / dt, r.n_iter_h.double().mean().item()))
try:
    print("kernel ms", sol.last_kernel_ms())
except Exception as e:
    print(e)
