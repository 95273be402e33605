"""Iteration-count histogram of the config-3 workload (65,536 scenarios) + per-batch-size kernel times."""
import os, sys, tempfile
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (R, os.path.join(R, "tests"), os.path.join(R, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch, helpers
from harmonic_power_flow_b200 import BatchSolver, scenarios
net, st, _ = helpers.packed_from_files("net3", 25, True, tempfile.mkdtemp())
sol = BatchSolver(net)
P, Q, I_N = scenarios.make_batch(net, 65536, "tight")
dP, dQ, dI = sol.prepare(P, Q, I_N)
r = sol.solve(dP, dQ, dI); torch.cuda.synchronize()
it = r.n_iter_h.cpu().numpy()
print("hist", np.bincount(it).tolist())
sol.set_profiling(True)
for B in (2048, 4096, 8192, 16384, 32768, 65536):
    best = [1e9, 1e9]
    for _ in range(5):
        rr = sol.solve(dP[:, :B].contiguous(), dQ[:, :B].contiguous(), dI[:, :, :B].contiguous())
        k = sol.last_kernel_ms(); best = [min(best[0], k[0]), min(best[1], k[1])]
    print("B=%d fund+wn %.4f harm %.4f ms max it %d" % (B, best[0], best[1], int(it[:B].max())))
